#!/bin/bash
# round 2, GPU call 22 (one B200): compute-sanitizer racecheck + synccheck over the queued necklace kernel (shared-memory
# queue, warp-level atomics) and the reduced paths on small cases
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 50 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_aux.py -m gpu -q -x \
    -k "staged-10-0 or (4x4_momentum and staged) or hash_and_buckets" > gpurun_out/r2c22_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -n "SUMMARY\|hazard\|passed\|failed\|Barrier error\|Divergent" gpurun_out/r2c22_$tool.log | head -8
done
