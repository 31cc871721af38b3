#!/bin/bash
# round 2, GPU call 11 (eight B200): refined partition + side-stream pack/fence; final N=8 / N=4 lines
mkdir -p gpurun_out
run() {
  local name=$1 n=$2; shift; shift
  ( time timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c11_$name.json 2> gpurun_out/r2c11_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c11_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), 'phases', d['details'].get('phases_run_back_to_back_ms'), 'GBps', d['details'].get('halo_GBps_per_rank'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), d.get('lanczos',{}).get('lowest_ritz'), 'chk', d['details']['checksum_x_dot_Hx'])
    t=d.get('tri6x6')
    if t: print('   tri6x6 free', round(t['matrix_free']['ms_per_matvec'],2), 'csr', round(t['cached_csr']['ms_per_matvec'],3), t['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c11_$name.err
}
BARGS="" run n8_full 8 X=1
BARGS="--no-extras --no-e2e" run n8_norefine 8 EDCUDA_SHARD_REFINE=0
BARGS="" run n4_full 4 X=1
