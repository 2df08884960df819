#!/usr/bin/env python
"""bench.py — images/sec of feather::Net::Forward on synthetic 224x224x3 input (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model vgg16|resnet50|mobilenet_v1|single_conv]
                    [--batch B_per_gpu] [--precision tf32x3|tf32] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

One "step" = one Forward over one batch.  Default workload = BASELINE.json configs[1]: VGG-16 fp32, batch 64 per
GPU (weak scaling: every rank runs its own 64-image shard; the only collective is the weight broadcast at load).
Prints ONE JSON line on rank 0:
  value         whole-job images/s with the input batches already resident in HBM (CUDA events, max over ranks)
  e2e           same metric through the public API with pinned HOST buffers: H2D of the batch and D2H of the
                softmax output inside the timed region, every step
  roofline      every instrumented launch of one eager Forward, timed with CUDA events on the Net's stream and grouped by
                kernel class; the object describes the class with the largest share of the step (achieved =
                algorithmic direct-conv FLOPs, or algorithmic bytes for HBM-bound classes, / summed duration, against
                MEASURED_PEAKS.json) and lists the others under other_kernels
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, bounded sample, rank 0 / N=1 only
`--impl reference` times only that CPU reference arm (no GPU work) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DEFAULT_BATCH = {"vgg16": 64, "resnet50": 128, "mobilenet_v1": 256, "single_conv": 64}
L2_BYTES = 126 * 1024 * 1024


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    # fallback stated by /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def model_files(model: str, rank: int, barrier) -> tuple[str, str, float, tuple]:
    """Rank 0 writes the synthetic ncnn .param/.bin once per box (random-init weights, fixed seed)."""
    from feathercnn_b200.tools import modelgen
    cache = Path(os.environ.get("FEATHER_BENCH_CACHE", "/tmp/feather_bench_models"))
    prefix = cache / f"{model}_seed0"
    meta = cache / f"{model}_seed0.json"
    if rank == 0 and not (meta.exists() and Path(str(prefix) + ".bin").exists()):
        cache.mkdir(parents=True, exist_ok=True)
        m = modelgen.ZOO[model]()
        m.save(prefix)
        meta.write_text(json.dumps({"flops": m.flops, "input": list(m.shape["data"])}))
    barrier()
    info = json.loads(meta.read_text())
    return str(prefix) + ".param", str(prefix) + ".bin", float(info["flops"]), tuple(info["input"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe).

    nvidia-smi needs ~0.1-0.5 s to deliver its first row, longer than a short timed region, so the poller is started
    before the warm-up and only the rows received inside the marked windows (device-resident leg and end-to-end leg,
    both under the same load) are used."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows: list[tuple[float, list[str]]] = []
        self.windows: list[list[float]] = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.windows.append([time.perf_counter(), float("inf")])

    def end(self):
        if self.windows:
            self.windows[-1][1] = time.perf_counter()

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(slack: float):
            sm, mx, reasons = [], [], set()
            for t, r in self.rows:
                if len(r) < 8 or not any(w0 - slack <= t <= w1 + slack for w0, w1 in self.windows):
                    continue
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                except ValueError:
                    continue
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, reasons

        sm, mx, reasons = collect(0.0)
        how = "inside the timed regions"
        if not sm:  # a region shorter than the polling period: take the rows right around it
            sm, mx, reasons = collect(0.12)
            how = "within 120 ms of the timed regions"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": how,
                "window_s": round(sum(min(w1, time.perf_counter()) - w0 for w0, w1 in self.windows), 3)}


def cpu_reference_arm(param: str, binf: str, batch_hint: int, warmup: int, steps: int) -> dict:
    from oracle import cpu_bench
    procs = cpu_bench.physical_cores()
    r = cpu_bench.run(param, binf, procs=procs, warmup=max(1, warmup), iters=max(1, steps))
    r["sample"] = (f"{r['procs']} single-thread processes x ({max(1, warmup)} warm-up + {max(1, steps)} timed) "
                   f"whole-net Forward of the same model, batch 1 each (the reference has no batch dimension)")
    return r


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    param, binf, flops, in_shape = model_files(args.model, 0, lambda: None)
    # bounded sample: each "step" is one Forward per worker process
    r = cpu_reference_arm(param, binf, args.batch, min(args.warmup, 1), min(args.steps, 3))
    line = {
        "impl": "reference", "metric": "images/sec", "value": r["images_per_s"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["sec_per_forward_mean"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model}_b{args.batch}_{in_shape[1]}x{in_shape[2]}", "batch_per_gpu": args.batch,
                   "impl_detail": "unmodified FeatherCNN AVX build (oracle/_ref); the reference has no batch dimension "
                                  "(src/blob.cpp:73) and is race-free only at 1 thread (src/net.cpp:38), so the batch is "
                                  "run as independent single-image Forwards, one single-thread process per physical core",
                   "timed_forwards_per_process": min(args.steps, 3)},
        "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["procs"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="vgg16", choices=["vgg16", "resnet50", "mobilenet_v1", "single_conv"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: BASELINE.json config)")
    ap.add_argument("--precision", default="tf32x3", choices=["tf32x3", "tf32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2-chunk-mb", type=float, default=None)
    ap.add_argument("--lean", action="store_true", help="profiling aid: device-resident leg only (no e2e / roofline / cpu legs)")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = DEFAULT_BATCH[args.model]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from feathercnn_b200 import booster
    from feathercnn_b200 import dist as fdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    booster.set_precision(booster.PRECISION_TF32 if args.precision == "tf32" else booster.PRECISION_TF32X3)
    if args.l2_chunk_mb is not None:
        booster.set_l2_chunk_bytes(int(args.l2_chunk_mb * 1024 * 1024))

    param, binf, flops_per_image, in_shape = model_files(args.model, rank, barrier)
    t_load = time.perf_counter()
    net = fdist.load_net_distributed(param, binf, local_rank, fusion=not args.no_fusion, cuda_graph=not args.no_graph)
    stream = torch.cuda.Stream(device=dev)
    net.SetStream(stream.cuda_stream)
    t_load = time.perf_counter() - t_load

    from feathercnn_b200.tools import modelgen
    B = args.batch
    batch_bytes = B * int(np.prod(in_shape)) * 4
    n_rot = max(2, -(-int(1.25 * L2_BYTES) // batch_bytes))  # rotating inputs larger than L2 in total
    n_rot = min(n_rot, 8)
    host_batches = []
    for r in range(n_rot):
        hb = torch.empty((B,) + tuple(in_shape), dtype=torch.float32).pin_memory()
        base = np.stack([modelgen.synthetic_input(in_shape, (rank * n_rot + r) * 4 + i) for i in range(min(B, 4))])
        hb.copy_(torch.from_numpy(np.resize(base, (B,) + tuple(in_shape))))
        host_batches.append(hb)
    dev_batches = [hb.to(dev) for hb in host_batches]
    out_name = "prob" if "prob" in net.BlobNames() else sorted(net.BlobNames())[-1]
    torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg ("value") -------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    with torch.cuda.stream(stream):
        for i in range(args.warmup + n_rot):  # eager pass, then one graph capture per rotating buffer
            d = dev_batches[i % n_rot]
            net.FeedInputDevice(d.data_ptr(), tuple(d.shape))
            net.Forward()
        net.Synchronize()
        launches = net.launches_per_forward
        barrier()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            d = dev_batches[i % n_rot]
            net.FeedInputDevice(d.data_ptr(), tuple(d.shape))
            net.Forward()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize(dev)
        barrier()
        if sampler:
            sampler.end()
        ms_dev = max_over_ranks(e0.elapsed_time(e1))

        if args.lean:
            if sampler:
                sampler.stop()
            if rank == 0:
                print(json.dumps({"lean": True, "value": world * B * args.steps / (ms_dev * 1e-3),
                                  "ms_per_step": ms_dev / args.steps, "launches_per_step": int(launches)}), flush=True)
            if world > 1:
                dist.destroy_process_group()
            return

        # ---- end-to-end leg: pinned host input -> H2D -> Forward -> D2H of the result, every step -----------
        n, c, h, w = net.BlobShape(out_name)
        host_out = torch.empty((n, c, h, w), dtype=torch.float32).pin_memory()
        for i in range(max(2, args.warmup)):
            net.ForwardBatchHostPtr(host_batches[i % n_rot].data_ptr(), B)
            net.ExtractInto(out_name, host_out.data_ptr())
        barrier()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.begin()
        e0.record(stream)
        t_wall = time.perf_counter()
        for i in range(args.steps):
            net.ForwardBatchHostPtr(host_batches[i % n_rot].data_ptr(), B)
            net.ExtractInto(out_name, host_out.data_ptr())  # D2H + stream sync: the caller holds the result
        e1.record(stream)
        e1.synchronize()
        t_wall = time.perf_counter() - t_wall
        if sampler:
            sampler.end()
        clocks = sampler.stop() if sampler else None
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        d2h_bytes = host_out.numel() * 4

        # ---- roofline leg: per-launch CUDA events around every TensorGEMM of one eager Forward ---------------
        roof = None
        if rank == 0:
            try:
                import ctypes
                lib = booster.fcuda()
                net._lib.fnet_set_cuda_graph(net._h, 0)
                net.FeedInputDevice(dev_batches[0].data_ptr(), tuple(dev_batches[0].shape))
                net.Forward(); net.Synchronize()
                lib.fcuda_profile_tensor_gemm(1)
                reps = 2
                for i in range(reps):
                    d = dev_batches[i % n_rot]
                    net.FeedInputDevice(d.data_ptr(), tuple(d.shape))
                    net.Forward()
                net.Synchronize()
                peaks, how = measured_peaks()
                # kernels timed inside a long step -> the sustained figures; fallbacks = /opt/skills/guides/B200_PROFILING.md
                tf_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
                hbm_peak = float(peaks.get("hbm_gbs", 6500.0))
                names = ["tensor_gemm_ts_kernel (tcgen05 kind::tf32, Winograd/im2col/FC GEMM)",
                         "conv_igemm_kernel (tcgen05 kind::tf32 implicit-GEMM conv)", "wino_input_kernel",
                         "wino_output_kernel", "pooling_kernel", "depthwise kernels", "element-wise kernels"]
                classes = []
                for kind, name in enumerate(names):
                    ms_t, af, mf, ab, nl = (ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double(),
                                            ctypes.c_longlong())
                    lib.fcuda_profile_collect_kind(kind, ctypes.byref(ms_t), ctypes.byref(af), ctypes.byref(mf),
                                                   ctypes.byref(ab), ctypes.byref(nl))
                    if nl.value == 0 or ms_t.value <= 0:
                        continue
                    sec = ms_t.value * 1e-3
                    tensor = kind <= 1
                    ach = af.value / sec / 1e12 if tensor else ab.value / sec / 1e9
                    peak = tf_peak if tensor else hbm_peak
                    c = {"kernel": name, "bound": "tensor" if tensor else "hbm", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak,
                         "launches_per_step": nl.value / reps, "avg_launch_us": 1e3 * ms_t.value / nl.value,
                         "share_of_step": (ms_t.value / reps) / (ms_dev / args.steps)}
                    if tensor:
                        c["algorithmic_gflop_per_launch"] = af.value / nl.value / 1e9
                        c["algorithmic_gb_per_launch"] = ab.value / nl.value / 1e9
                        c["hbm_gbps_at_algorithmic_bytes"] = ab.value / sec / 1e9
                        c["tensor_pipe_tflops_issued"] = mf.value / sec / 1e12  # 3 MMAs per product in 3xTF32 mode
                        c["tensor_pipe_frac_of_tf32_peak"] = c["tensor_pipe_tflops_issued"] / (tf_peak / 2.0)
                    else:
                        c["algorithmic_gb_per_launch"] = ab.value / nl.value / 1e9
                    classes.append(c)
                lib.fcuda_profile_tensor_gemm(0)
                if classes:
                    classes.sort(key=lambda c: -c["share_of_step"])
                    roof = dict(classes[0])
                    roof["peak_source"] = (f"{how}: " + ("bf16_tflops_sustained" if roof["bound"] == "tensor" else
                                                        "hbm_gbs") + " (kernel timed inside a long step)")
                    tr = ROOT / "profiles" / "r01_kernel_traffic.json"
                    roof["traffic"] = None
                    if tr.exists():
                        try:
                            roof["traffic"] = json.loads(tr.read_text()).get(args.model, {}).get(roof["kernel"].split()[0])
                        except Exception:
                            pass
                    roof["other_kernels"] = classes[1:]
            except Exception as e:  # the roofline leg explains the number, it must never cost the bench line
                roof = {"error": f"{type(e).__name__}: {e}"[:300]}

    total_images = world * B * args.steps
    value = total_images / (ms_dev * 1e-3)
    e2e_value = total_images / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_arm(param, binf, B, 1, 3)
            cpu = {"value": r["images_per_s"], "unit": "images/s", "cores": r["procs"], "kind": r["kind"],
                   "sample": r["sample"], "sec_per_forward_per_core": r["sec_per_forward_mean"]}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "images/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}

    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "tf32x3" else "tf32", "data": "synthetic",
        "config": {"workload": f"{args.model}_b{B}_{in_shape[1]}x{in_shape[2]}", "batch_per_gpu": B,
                   "global_batch": B * world, "parallelism": f"batch-shard x{world} (weights broadcast once over NCCL)",
                   "tensor_core_mode": "3xTF32 split (fp32-equivalent)" if args.precision == "tf32x3" else "TF32",
                   "algorithms": "tuned SelectAlgo (reference rule, then Winograd -> implicit GEMM when IC,OC <= 128 and OW >= 28, "
                                 "im2col -> implicit GEMM): Winograd F(6,3)+TensorGEMM / SGECONV implicit GEMM / depthwise",
                   "fusion": not args.no_fusion, "cuda_graph": not args.no_graph,
                   "l2_policy": f"{n_rot} rotating input batches ({n_rot * batch_bytes / 2**20:.0f} MiB > L2); "
                                f"activations per step far exceed the 126 MB L2",
                   "gflop_per_image": flops_per_image / 1e9, "load_s": round(t_load, 3)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": batch_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "wall_ms_per_step": 1e3 * t_wall / args.steps},
        "gpu_launches": int(launches) * args.steps,
        "launches_per_step": int(launches),
        "algorithmic_tflops": value * flops_per_image / 1e12,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
