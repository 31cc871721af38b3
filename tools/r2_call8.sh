#!/bin/bash
# round 2, GPU call 8 (eight B200): rotated peer order in the halo exchange; pull-stream count; final N=8 / N=4 lines
mkdir -p gpurun_out
run() {
  local name=$1 n=$2; shift; shift
  ( time timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c8_$name.json 2> gpurun_out/r2c8_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c8_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), 'phases', d['details'].get('phases_run_back_to_back_ms'), 'GBps', d['details'].get('halo_GBps_per_rank'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), d.get('lanczos',{}).get('lowest_ritz'), 'chk', d['details']['checksum_x_dot_Hx'])
    t=d.get('tri6x6')
    if t: print('   tri6x6 free', round(t['matrix_free']['ms_per_matvec'],2), 'csr', round(t['cached_csr']['ms_per_matvec'],3), t['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c8_$name.err
}
BARGS="--no-extras --no-e2e" run n8_pull1 8 X=1
BARGS="--no-extras --no-e2e" run n8_pull2 8 EDCUDA_PULL_STREAMS=2
BARGS="--no-extras --no-e2e" run n8_pull4 8 EDCUDA_PULL_STREAMS=4
BARGS="--no-extras --no-e2e --exchange cepush" run n8_cepush 8 X=1
BARGS="" run n8_full 8 X=1
BARGS="" run n4_full 4 X=1
BARGS="--no-extras --no-e2e" run n4_pull2 4 EDCUDA_PULL_STREAMS=2
