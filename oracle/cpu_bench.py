"""TEST INFRASTRUCTURE ONLY — CPU baseline timer used by bench.py (cpu_baseline leg and `--impl reference`).

Times the reference's own CPU implementation of the hot path on the host cores: P independent single-thread
worker processes (the reference is only race-free at 1 thread — /root/reference/src/net.cpp:38,
avx/winograd_kernels_F63.cpp:542-546 — so this is its best embarrassingly-parallel case), each
looping whole-net Forward on its own feather::Net from oracle/_ref (kind "reference"), or — if that library is
absent — on the NumPy/C restatement (kind "port").

Worker protocol: `python -m oracle.cpu_bench worker <param> <bin> <core> <warmup> <iters>` prints
"<seconds per forward>".  run() launches the workers, pins them to distinct cores and aggregates.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def physical_cores() -> int:
    try:
        pairs = set()
        phys = core = None
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("physical id"):
                phys = line.split(":")[1].strip()
            elif line.startswith("core id"):
                core = line.split(":")[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    pairs.add((phys, core))
                phys = core = None
        if pairs:
            return min(len(pairs), len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return max(1, len(os.sched_getaffinity(0)) // 2)


def _worker(param: str, binf: str, core: int, warmup: int, iters: int) -> None:
    try:
        allowed = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {allowed[core % len(allowed)]})
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, str(ROOT))
    from feathercnn_b200.tools import modelgen
    from oracle import oracle as O
    layers = O.parse_param(param)
    pd = layers[0]["params"]
    shape = (pd.get(2, 3), pd.get(1, 224), pd.get(0, 224))
    x = modelgen.synthetic_input(shape, core)
    if O.reference_available():
        net = O.ReferenceNet(param, binf)
        for _ in range(max(warmup - 1, 0)):
            net.forward(x)
        sec = net.time_forward(x, iters)  # one more warm-up (lazy Init) + `iters` timed Forwards
        kind = "reference"
    else:
        net = O.OracleNet(param, binf)
        for _ in range(max(warmup, 1)):
            net.forward(x)
        t0 = time.perf_counter()
        for _ in range(iters):
            net.forward(x)
        sec = (time.perf_counter() - t0) / iters
        kind = "port"
    print(f"{sec:.6f} {kind}", flush=True)


def run(param: str, binf: str, procs: int | None = None, warmup: int = 1, iters: int = 3) -> dict:
    """Aggregate images/s of `procs` single-thread workers (default: one per physical core)."""
    procs = procs or physical_cores()
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=str(ROOT))
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, "-m", "oracle.cpu_bench", "worker", str(param), str(binf), str(i),
                            str(warmup), str(iters)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env,
                           cwd=str(ROOT), text=True) for i in range(procs)]
    secs, kind = [], "reference"
    for p in ps:
        out, _ = p.communicate()
        try:
            s, kind = out.strip().split()[-2:]
            secs.append(float(s))
        except Exception:
            pass
    wall = time.perf_counter() - t0
    if not secs:
        raise RuntimeError("cpu baseline workers produced no timing")
    ips = float(sum(1.0 / s for s in secs))
    return {"images_per_s": ips, "procs": len(secs), "sec_per_forward_mean": float(np.mean(secs)),
            "sec_per_forward_max": float(np.max(secs)), "kind": kind, "iters": iters, "wall_s": wall}


if __name__ == "__main__":
    if len(sys.argv) >= 7 and sys.argv[1] == "worker":
        _worker(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))
    else:
        import json
        print(json.dumps(run(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)))
