#!/bin/bash
# round 2, GPU call 19 (one B200): x tile staged by a TMA bulk copy in k2_apply_u1: full GPU suite, bench with / without
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2c19_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2c19_pytest.log | tail -2
for v in bulk nobulk; do
  if [ $v = nobulk ]; then export EDCUDA_U1_NOBULK=1; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-e2e > gpurun_out/r2c19_$v.json 2> gpurun_out/r2c19_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c19_$v.json')); print('$v', d['ms_per_step'], d['roofline']['frac'], d['details']['checksum_x_dot_Hx'], d['clocks'])"
done
