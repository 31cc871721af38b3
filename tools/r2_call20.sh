#!/bin/bash
# round 2, GPU call 20 (two B200): final build through the multi-GPU bench path
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2c20_bench_n2.json 2> gpurun_out/r2c20_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c20_bench_n2.json'))
t=d['tri6x6']
print('N=2 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'lanczos', d['lanczos']['ms_per_step'], d['lanczos']['lowest_ritz'], d['details'].get('phases_run_back_to_back_ms'), 'chk', d['details']['checksum_x_dot_Hx'], 'clocks', d['clocks'])
print('  tri free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'], t['checksum_x_dot_Hx'])
PY
tail -3 gpurun_out/r2c20_bench_n2.err
