"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
Only launches after the warm-up (the last bench step) matter for shares; this prints both views."""
import csv
import re
import sys
from collections import defaultdict

rows = []
traffic = {}  # launch ID -> DRAM bytes (read + write), when the list was collected with the dram__bytes metrics too
ids = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name", "").startswith("dram__bytes"):
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "byte")
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        traffic[r["ID"]] = traffic.get(r["ID"], 0.0) + v
        continue
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    ids.append(r["ID"])
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    name = re.sub(r"^void ", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((name, val * scale))
if not rows:
    print("no launches parsed")
    sys.exit(0)


def table(sel, title):
    tot = sum(v for _, v in sel)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in sel:
        agg[n][0] += 1
        agg[n][1] += v
    print(f"== {title}: {len(sel)} launches, {tot / 1e3:.3f} ms total (cold-cache serialised times)")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / tot * 100:6.2f}%  {v / 1e3:9.3f} ms  {c:5d} x {v / c:9.1f} us  {n[:110]}")


table(rows, "all launches")
# the last quarter of the list is the timed step (1 eager step after 3+rot warm-ups and the e2e/roofline legs vary);
# use the final occurrence of the softmax kernel as the end of a step and the previous one as its start.
idx = [i for i, (n, _) in enumerate(rows) if "softmax_kernel" in n]
if len(idx) >= 2:
    table(rows[idx[-2] + 1: idx[-1] + 1], "one Forward (between the last two softmax launches)")
    if traffic:
        # measured DRAM traffic per launch of the step, per kernel: what bench.py reports as roofline.traffic
        import json
        agg = defaultdict(lambda: [0, 0.0])
        for i in range(idx[-2] + 1, idx[-1] + 1):
            k = rows[i][0].split("<")[0].split("::")[-1]
            agg[k][0] += 1
            agg[k][1] += traffic.get(ids[i], 0.0)
        print("== measured DRAM bytes (read+write) per launch, one Forward")
        out = {}
        for k, (c, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{b / c / 1e6:10.1f} MB x {c:3d}  {k}")
            out[k] = b / c
        if len(sys.argv) > 2:
            json.dump(out, open(sys.argv[2], "w"), indent=1)
