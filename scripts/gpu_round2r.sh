#!/bin/bash
# round 2r: ncu --set full with source of the implicit-GEMM conv (conv1_2, conv2_1, conv2_2 of VGG-16, BF16x3 mode)
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02r_*
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 13 -c 3 -o $O/r02r_vgg16_igemm -f \
    python bench.py --model vgg16 --steps 1 --warmup 3 --no-graph --lean > $O/r02r_ncu.stdout 2>&1
tail -3 $O/r02r_ncu.stdout
ncu -i $O/r02r_vgg16_igemm.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py > $O/r02r_vgg16_igemm_summary.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|issue_active|stalled" $O/r02r_vgg16_igemm_summary.txt | cut -c1-150
ncu -i $O/r02r_vgg16_igemm.ncu-rep --page source --csv > $O/r02r_vgg16_igemm_source.csv 2>/dev/null
ls -la $O/r02r_*; head -c 1500 $O/r02r_vgg16_igemm_source.csv
