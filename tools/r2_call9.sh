#!/bin/bash
# round 2, GPU call 9 (two B200): final-build validation: full GPU suite, N=1 and N=2 default bench lines, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2c9_pytest.log 2>&1
tail -5 gpurun_out/r2c9_pytest.log
( time EDCUDA_K6_TIMING=1 timeout 900 python bench.py ) > gpurun_out/r2c9_bench_n1.json 2> gpurun_out/r2c9_bench_n1.err
grep "K6 staged" gpurun_out/r2c9_bench_n1.err | tail -1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench_n1.json'))
t=d['tri6x6']
print('N=1 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'lanczos', d['lanczos']['ms_per_step'], d['lanczos']['lowest_ritz'], 'parity', d['parity_sample'], 'clocks', d['clocks'])
print('  tri free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'], 'assemble', t['cached_csr']['assemble_seconds'], 'sparse', d['sparse'])
print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2c9_bench_n2.json 2> gpurun_out/r2c9_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench_n2.json'))
t=d['tri6x6']
print('N=2 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'lanczos', d['lanczos']['ms_per_step'], d['lanczos']['lowest_ritz'], d['details'].get('phases_run_back_to_back_ms'), 'clocks', d['clocks'])
print('  tri free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'])
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r2c9_ref_n2.json 2> gpurun_out/r2c9_ref_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2c9_ref_n2.json')); print('ref arm under torchrun: cores', d['cpu_baseline']['cores'], 'value', d['value'], d['config'])"
