#!/bin/bash
# round 2, GPU call 13 (one B200): hash index of the reduced basis: full GPU suite + config 4 timing
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2c13_pytest.log 2>&1
tail -5 gpurun_out/r2c13_pytest.log
for h in 1 0; do
  ( time EDCUDA_RBASIS_HASH=$h EDCUDA_K6_TIMING=1 timeout 600 python bench.py --workload tri6x6_k0A1_sz0 --steps 5 ) > gpurun_out/r2c13_tri_hash$h.json 2> gpurun_out/r2c13_tri_hash$h.err
  grep "K6 staged" gpurun_out/r2c13_tri_hash$h.err | tail -1
  python - <<PY
import json
try:
    t=json.load(open('gpurun_out/r2c13_tri_hash$h.json'))['tri6x6']
    print('hash=$h free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'], 'assemble', t['cached_csr']['assemble_seconds'], 'setup', t['setup_seconds'], t['checksum_x_dot_Hx'])
except Exception as e:
    print('hash=$h FAILED', e)
PY
done
