#!/bin/bash
# round 2, GPU call 6 (eight B200): halo transports at N=8 / N=4 (reader pulls, owner SM pushes, owner copy-engine pushes)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q ) > gpurun_out/r2c6_pytest.log 2>&1
tail -3 gpurun_out/r2c6_pytest.log
run() {
  local name=$1 n=$2; shift; shift
  ( time timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c6_$name.json 2> gpurun_out/r2c6_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c6_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), 'phases', d['details'].get('phases_run_back_to_back_ms'), 'GBps', d['details'].get('halo_GBps_per_rank'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), d.get('lanczos',{}).get('lowest_ritz'), 'chk', d['details']['checksum_x_dot_Hx'])
    t=d.get('tri6x6')
    if t: print('   tri6x6 free', round(t['matrix_free']['ms_per_matvec'],2), 'csr', round(t['cached_csr']['ms_per_matvec'],3), t['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c6_$name.err
}
BARGS="--no-extras --no-e2e --exchange pull" run n8_pull 8 X=1
BARGS="--no-extras --no-e2e --exchange push" run n8_push64 8 X=1
BARGS="--no-extras --no-e2e --exchange push" run n8_push24 8 EDCUDA_PUSH_CTAS=24
BARGS="--no-extras --no-e2e --exchange cepush" run n8_cepush 8 X=1
BARGS="--no-extras --no-e2e --exchange cepush --chunks 4" run n8_cepush_c4 8 X=1
BARGS="--exchange push" run n8_full_push 8 X=1
BARGS="--exchange cepush" run n8_full_cepush 8 X=1
BARGS="--no-extras --no-e2e --exchange push" run n4_push 4 X=1
BARGS="--no-extras --no-e2e --exchange cepush" run n4_cepush 4 X=1
( time EDCUDA_K6_TIMING=1 timeout 600 python bench.py --workload tri6x6_k0A1_sz0 --steps 3 ) > gpurun_out/r2c6_tri_n1.json 2> gpurun_out/r2c6_tri_n1.err
grep "K6 staged" gpurun_out/r2c6_tri_n1.err | tail -1
python -c "
import json; d=json.load(open('gpurun_out/r2c6_tri_n1.json'))['tri6x6']; print('tri N=1 free', d['matrix_free']['ms_per_matvec'], 'csr', d['cached_csr']['ms_per_matvec'], d['cached_csr']['assemble_seconds'], d['checksum_x_dot_Hx'])"
