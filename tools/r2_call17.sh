#!/bin/bash
# round 2, GPU call 17 (one B200): queued two-pass necklace kernel: reduced-path tests + config 4 timing, queued / one-pass
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_aux.py -m gpu -q -k "tri or reduced or symmetry or square or config2 or config4 or sector" ) > gpurun_out/r2c17_pytest.log 2>&1
tail -3 gpurun_out/r2c17_pytest.log
( time EDCUDA_K6_TIMING=1 timeout 600 python bench.py --workload tri6x6_k0A1_sz0 --steps 5 ) > gpurun_out/r2c17_tri_queued.json 2> gpurun_out/r2c17_tri_queued.err
grep "K6 staged" gpurun_out/r2c17_tri_queued.err | tail -1
python - <<PY
import json
t=json.load(open('gpurun_out/r2c17_tri_queued.json'))['tri6x6']
print('queued free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'], t['checksum_x_dot_Hx'])
PY
