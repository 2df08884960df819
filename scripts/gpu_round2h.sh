#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02h_*
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02h_pytest_gpu.log
grep -E "^FAILED|passed|failed|rc=" $O/r02h_pytest_gpu.log | tail -20
timeout 300 python bench.py --lean --model vgg16 > $O/r02h_lean.log 2>&1; cat $O/r02h_lean.log | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel --launch-count 4 \
  -o $O/r02h_vgg16_igemm -f python bench.py --lean --no-graph --model vgg16 --steps 1 --warmup 3 > $O/r02h_ncu_igemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tensor_gemm_ts_kernel --launch-count 3 \
  -o $O/r02h_vgg16_gemm -f python bench.py --lean --no-graph --model vgg16 --steps 1 --warmup 3 > $O/r02h_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wino_(in|out)put_kernel" --launch-skip 2 --launch-count 2 \
  -o $O/r02h_vgg16_wino -f python bench.py --lean --no-graph --model vgg16 --steps 1 --warmup 3 > $O/r02h_ncu_wino.log 2>&1
ls -la $O/r02h_*
