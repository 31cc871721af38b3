#!/bin/bash
# round 2, GPU call 18 (one B200): final build: full GPU suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2c18_pytest.log 2>&1
tail -4 gpurun_out/r2c18_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/r2c18_bench_n1.json 2> gpurun_out/r2c18_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c18_bench_n1.json'))
t=d['tri6x6']
print('N=1 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'lanczos', d['lanczos']['ms_per_step'], d['lanczos']['lowest_ritz'], 'parity', d['parity_sample']['max_rel_err'], 'clocks', d['clocks'])
print('  tri free', t['matrix_free']['ms_per_matvec'], 'csr', t['cached_csr']['ms_per_matvec'], 'assemble', t['cached_csr']['assemble_seconds'])
print('  sparse', d['sparse'])
print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
