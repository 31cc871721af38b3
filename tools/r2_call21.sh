#!/bin/bash
# round 2, GPU call 21 (one B200): compute-sanitizer memcheck over the kernels added last (bulk-copy staging, queued necklace
# kernel, hash look-up, sharded launches) on small cases
mkdir -p gpurun_out
timeout 135 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_aux.py tests/test_gpu_sharded.py -m gpu -q -x \
  -k "staged-10-0 or (4x4_momentum and staged) or 16-8-square or 13-6-open or hash_and_buckets" > gpurun_out/r2c21_memcheck.log 2>&1
echo "rc=$?"
grep -n "ERROR SUMMARY\|Invalid\|passed\|failed\|out of bounds\|misaligned" gpurun_out/r2c21_memcheck.log | head -20
