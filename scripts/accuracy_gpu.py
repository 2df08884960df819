"""Whole-net accuracy of both tensor-core modes against the reference build (GPU box)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feathercnn_b200 import booster
from feathercnn_b200.net import Net
from feathercnn_b200.tools import modelgen
from oracle import oracle as O

for name in sys.argv[1:] or ["vgg16", "resnet50", "mobilenet_v1"]:
    m = modelgen.ZOO[name]()
    param, binf = m.save(f"/tmp/acc_{name}")
    x = modelgen.synthetic_input(m.shape["data"], 0)
    cpu = O.ReferenceNet(param, binf) if O.reference_available() else O.OracleNet(param, binf)
    cpu.forward(x)
    for mode, label in ((booster.PRECISION_FP32_SPLIT, "fp32split"), (booster.PRECISION_TF32X3, "3xTF32"), (booster.PRECISION_TF32, "TF32")):
        booster.set_precision(mode)
        net = Net()
        net.LoadParam(param); net.LoadWeights(binf)
        net.Forward(x[None])
        errs = {}
        for b in sorted(m.blobs):
            ref = cpu.extract(b)
            got = net.Extract(b)[0]
            errs[b] = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
        worst = max(errs, key=errs.get)
        print(f"{name:14s} {label:9s} worst blob {worst:28s} {errs[worst]:.3e}   prob {errs.get('prob', float('nan')):.3e}   "
              f"median {np.median(list(errs.values())):.3e}", flush=True)
    booster.set_precision(booster.PRECISION_FP32_SPLIT)
