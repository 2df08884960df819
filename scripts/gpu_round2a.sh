#!/bin/bash
# first GPU pass of round 2: micro-probe, parity tests, default bench line (3 workloads)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02a_box.txt; nproc >> gpurun_out/r02a_box.txt
cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02a_box.txt 2>&1
timeout 60 build/tma_shift > gpurun_out/r02a_tma_shift.txt 2>&1; echo "tma_shift rc=$?" >> gpurun_out/r02a_tma_shift.txt
timeout 120 build/gemm_selftest > gpurun_out/r02a_gemm_selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/r02a_gemm_selftest.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_tma_shift.txt; tail -2 gpurun_out/r02a_gemm_selftest.log; tail -15 gpurun_out/r02a_pytest_gpu.log; tail -5 gpurun_out/r02a_bench.err; head -c 1500 gpurun_out/r02a_bench.json
