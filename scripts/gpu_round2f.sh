#!/bin/bash
# ncu --set full on the implicit-GEMM conv kernels and the Winograd TensorGEMM of one VGG-16 b64 Forward
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02f_*
export FCUDA_GEMM_CLUSTER=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel --launch-count 4 \
  -o $O/r02f_vgg16_igemm -f python bench.py --lean --no-graph --model vgg16 --steps 1 --warmup 3 > $O/r02f_ncu_igemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tensor_gemm_ts_kernel --launch-count 3 \
  -o $O/r02f_vgg16_gemm -f python bench.py --lean --no-graph --model vgg16 --steps 1 --warmup 3 > $O/r02f_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wino_ --launch-skip 2 --launch-count 2 \
  -o $O/r02f_vgg16_wino -f python bench.py --lean --no-graph --model vgg16 --steps 1 --warmup 3 > $O/r02f_ncu_wino.log 2>&1
ls -la $O/r02f_*; tail -3 $O/r02f_ncu_igemm.log
# per-class roofline of the full bench (eager profile leg) for slab on / off
FCUDA_IGEMM_SLAB=0 timeout 600 python bench.py --model vgg16 --no-cpu-baseline > $O/r02f_bench_slab0.json 2>$O/r02f_bench_slab0.err
FCUDA_IGEMM_SLAB=1 timeout 600 python bench.py --model vgg16 --no-cpu-baseline > $O/r02f_bench_slab1.json 2>$O/r02f_bench_slab1.err
