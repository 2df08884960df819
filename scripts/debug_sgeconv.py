"""Debug aid (GPU box): runs SGECONV cases in separate subprocesses so a trapped kernel cannot poison the next case."""
import subprocess
import sys

CASES = {
    # name: (oc, ic, h, w, k, pad, batch, precision)
    "pw_invalid_boxes": (64, 32, 8, 8, 1, 0, 1, 0),
    "pw_full_boxes": (64, 32, 8, 16, 1, 0, 1, 0),
    "pw_tf32": (64, 32, 8, 16, 1, 0, 1, 1),
    "k3x1_h_shift": (64, 32, 8, 32, (3, 1), (0, 1), 1, 0),   # (kh, kw), (pad_w, pad_h): shifts along H only
    "k1x3_w_shift": (64, 32, 8, 32, (1, 3), (1, 0), 1, 0),   # shifts along W: unaligned inner TMA coordinate
    "k3_pad0": (64, 32, 8, 32, 3, 0, 1, 0),
    "k3_pad1_exact": (64, 32, 4, 32, 3, 1, 1, 0),
    "k3_pad1_tf32": (64, 32, 4, 32, 3, 1, 1, 1),
    "k3_pad1_batch": (48, 32, 21, 28, 3, 1, 2, 0),
    "k3_vgg": (64, 64, 56, 56, 3, 1, 2, 0),
}

if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np
    import torch
    sys.path.insert(0, ".")
    from feathercnn_b200 import booster
    oc, ic, h, w, k, pad, batch, prec = CASES[sys.argv[2]]
    booster.set_precision(prec)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((batch, ic, h, w), device="cuda", generator=g) - 0.5
    b = torch.rand(oc, device="cuda", generator=g) - 0.5
    if not isinstance(k, tuple):
        wt = torch.randn((oc, ic, k, k), device="cuda", generator=g) * 0.05
    if isinstance(k, tuple):
        kh, kw = k
        pw, ph = pad
        wt = torch.randn((oc, ic, kh, kw), device="cuda", generator=g) * 0.05
        p = booster.ConvParam.make(oc, ic, h, w, kh, kw, pad_lbrt=(pw, ph, pw, ph), bias=True, relu=False)
        tpad = (ph, pw)
    else:
        p = booster.ConvParam.make(oc, ic, h, w, k, pad=pad, bias=True, relu=False)
        tpad = pad
    out, algo = booster.conv_forward(p, x, wt, b, algo=booster.SGECONV)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=tpad).float()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    print(f"{sys.argv[2]:18s} rel_err {err:.3e}  out[0,0,0,:4] {out[0, 0, 0, :4].tolist()} ref {ref[0, 0, 0, :4].tolist()}")
else:
    for name in (sys.argv[1:] or CASES):
        r = subprocess.run([sys.executable, __file__, "one", name], capture_output=True, text=True, timeout=120)
        tail = (r.stdout.strip().splitlines() or [""])[-1]
        err = [l for l in r.stderr.splitlines() if "rror" in l or "timed out" in l][-2:]
        print(name, "rc", r.returncode, "|", tail, "|", " ; ".join(err)[:300], flush=True)
        for l in r.stdout.splitlines():
            if "timed out" in l:
                print("   ", l)
