"""Forward the same batch several times and list the blobs that are not bit-identical run to run (first one = culprit)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from feathercnn_b200.net import Net  # noqa: E402
from feathercnn_b200.tools import modelgen  # noqa: E402

model, batch = sys.argv[1], int(sys.argv[2])
fusion = len(sys.argv) < 4 or sys.argv[3] != "nofusion"
d = Path("/tmp/detprobe")
d.mkdir(exist_ok=True)
m = modelgen.ZOO[model]()
param, binf = m.save(d / model)
x = np.stack([modelgen.synthetic_input(m.shape["data"], i) for i in range(batch)])
net = Net(fusion=fusion)
net.LoadParam(param)
net.LoadWeights(binf)
order = [l.split()[1] for l in Path(param).read_text().splitlines()[2:]]
runs = []
for r in range(3):
    net.Forward(x)
    got = {}
    for b in net.BlobNames():
        try:
            got[b] = net.Extract(b)
        except Exception:
            pass
    runs.append(got)
bad = []
for b in order:
    if b in runs[0] and any(not np.array_equal(runs[0][b], r[b]) for r in runs[1:]):
        a, c = runs[0][b], runs[1][b]
        bad.append((b, int((a != c).sum()), a.size, float(np.abs(a - c).max() / max(np.abs(a).max(), 1e-30))))
print(f"{model} b{batch} fusion={fusion}: {len(bad)} of {len(runs[0])} blobs differ run to run")
for b in bad[:12]:
    print("   first differing blobs:", b)
