#!/bin/bash
# round 2, GPU call 12 (two B200): windowed all-gather under the column-blocked cached SpMV: tests, config 4 at N=2 on / off
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --durations=5 ) > gpurun_out/r2c12_pytest.log 2>&1
tail -4 gpurun_out/r2c12_pytest.log
for win in 1 0; do
  ( time EDCUDA_SHARD_WINDOWS=$win timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$win bench.py --gpus 2 --steps 10 --warmup 3 --workload tri6x6_k0A1_sz0 ) > gpurun_out/r2c12_tri_n2_win$win.json 2> gpurun_out/r2c12_tri_n2_win$win.err
  python - <<PY
import json
try:
    t=json.load(open('gpurun_out/r2c12_tri_n2_win$win.json'))['tri6x6']
    print('windows=$win free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'], t['checksum_x_dot_Hx'])
except Exception as e:
    print('windows=$win FAILED', e)
PY
  tail -3 gpurun_out/r2c12_tri_n2_win$win.err
done
