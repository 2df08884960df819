#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02d_*
for s in 0 1; do for i in 1 2; do
  FCUDA_IGEMM_SLAB=$s FCUDA_IGEMM_ISSUERS=$i timeout 120 python scripts/debug_slab.py >> $O/r02d_debug.log 2>&1; echo "rc=$?" >> $O/r02d_debug.log
done; done
cat $O/r02d_debug.log
FCUDA_IGEMM_SLAB=1 FCUDA_IGEMM_ISSUERS=2 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/debug_slab.py 2 > $O/r02d_memcheck.log 2>&1
grep -v "^SLAB" $O/r02d_memcheck.log | head -60
