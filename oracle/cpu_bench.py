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


def cgroup_cpu_limit() -> float | None:
    """CPU quota of this container in cores (cgroup v2 cpu.max, v1 cfs_quota/cfs_period), or None when unlimited.
    A 1-GPU lease of a 64-core host may be capped at a fraction of it: round 1 sized the worker pool by
    /proc/cpuinfo alone and ran 4x oversubscribed there."""
    try:
        txt = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if txt and txt[0] != "max":
            return float(txt[0]) / float(txt[1])
        if txt:
            return None
    except Exception:
        pass
    try:
        q = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
        per = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
        if q > 0 and per > 0:
            return q / per
    except Exception:
        pass
    return None


def usable_cores() -> tuple[int, dict]:
    """Worker count for the embarrassingly-parallel CPU arm: physical cores, bounded by the affinity mask and by the
    container's CPU quota.  Returns (P, how) with the inputs of the decision for the bench record."""
    phys = physical_cores()
    quota = cgroup_cpu_limit()
    p = phys
    if quota is not None:
        p = max(1, min(phys, int(quota + 1e-6)))
    return p, {"physical_cores": phys, "affinity": len(os.sched_getaffinity(0)),
               "cgroup_quota_cores": None if quota is None else round(quota, 2)}


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
    """Aggregate images/s of `procs` single-thread workers (default: one per usable core, see usable_cores)."""
    procs = procs or usable_cores()[0]
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


def _dump(param: str, binf: str, index: int, out_npz: str) -> None:
    """Parity gate of bench.py: every blob of one reference Forward on synthetic image `index` -> npz."""
    sys.path.insert(0, str(ROOT))
    from feathercnn_b200.tools import modelgen
    from oracle import oracle as O
    layers = O.parse_param(param)
    pd = layers[0]["params"]
    shape = (pd.get(2, 3), pd.get(1, 224), pd.get(0, 224))
    x = modelgen.synthetic_input(shape, index)
    net = O.ReferenceNet(param, binf) if O.reference_available() else O.OracleNet(param, binf)
    net.forward(x)
    blobs = {}
    for layer in layers:
        for top in layer["tops"]:
            try:
                blobs[top] = net.extract(top)
            except Exception:
                pass
    np.savez(out_npz, __kind__=np.array("reference" if O.reference_available() else "port"), **blobs)


def dump_blobs(param: str, binf: str, index: int) -> tuple[dict, str]:
    """Runs _dump in a child process (the reference prints on its hot path, net.cpp:84,104,209) and loads the result."""
    import tempfile
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=str(ROOT))
    with tempfile.TemporaryDirectory() as d:
        out = str(Path(d) / "blobs.npz")
        subprocess.run([sys.executable, "-m", "oracle.cpu_bench", "dump", str(param), str(binf), str(index), out],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env, cwd=str(ROOT), check=True)
        z = np.load(out)
        kind = str(z["__kind__"])
        return {k: z[k] for k in z.files if k != "__kind__"}, kind


if __name__ == "__main__":
    if len(sys.argv) >= 7 and sys.argv[1] == "worker":
        _worker(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))
    elif len(sys.argv) >= 6 and sys.argv[1] == "dump":
        _dump(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
    else:
        import json
        print(json.dumps(run(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)))
