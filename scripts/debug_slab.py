"""SGECONV variants (slab producer on/off, 1 or 2 issuers) against the CUDA-core direct kernel on multi-tile shapes."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from feathercnn_b200 import booster  # noqa: E402

shapes = [(64, 64, 112, 112, 1), (64, 64, 224, 224, 1), (64, 64, 224, 224, 3), (128, 64, 112, 112, 4), (128, 128, 112, 112, 4),
          (64, 64, 56, 56, 16), (32, 32, 120, 100, 3), (64, 32, 60, 75, 5)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
# (oc, ic, h, w, n, k, stride, pad): generic (non-slab) producers as well
extra = [(256, 64, 56, 56, 8, 1, 1, 0), (64, 256, 56, 56, 8, 1, 1, 0), (128, 64, 57, 57, 4, 3, 2, 1), (64, 3, 224, 224, 2, 3, 1, 1),
         (128, 128, 28, 28, 16, 3, 1, 1)]
shapes = [s + (3, 1, 1) for s in shapes] + (extra if len(sys.argv) <= 1 else [])
rng = np.random.default_rng(0)
for oc, ic, h, w, n, k, stride, pad in shapes:
    x = torch.from_numpy(rng.uniform(-0.5, 0.5, (n, ic, h, w)).astype(np.float32)).cuda()
    wt = torch.from_numpy((rng.standard_normal((oc, ic, k, k)) * np.sqrt(2.0 / (ic * k * k))).astype(np.float32)).cuda()
    b = torch.from_numpy(rng.uniform(-0.1, 0.1, oc).astype(np.float32)).cuda()
    p = booster.ConvParam.make(oc, ic, h, w, k, stride=stride, pad=pad, relu=True)
    ref, _ = booster.conv_forward(p, x, wt, b, algo=booster.NAIVE)
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        out, _ = booster.conv_forward(p, x, wt, b, algo=booster.SGECONV)
        torch.cuda.synchronize()
        outs.append(out)
    err = float((outs[0] - ref).abs().max() / ref.abs().max())
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    tiles = n * ((h + 3) // 4) * ((w + 31) // 32)
    print(f"SLAB={os.environ.get('FCUDA_IGEMM_SLAB', '1')} ISSUERS={os.environ.get('FCUDA_IGEMM_ISSUERS', '2')} "
          f"CG={os.environ.get('FCUDA_IGEMM_CG', '-')} {ic}->{oc} {h}x{w} k{k}s{stride} b{n} tiles~{tiles}: rel_err {err:.2e} "
          f"deterministic={same}", flush=True)
