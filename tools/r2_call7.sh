#!/bin/bash
# round 2, GPU call 7 (two B200): NCCL send/recv halo transport: tests + N=2 bench vs pulls
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_aux.py -m gpu -q ) > gpurun_out/r2c7_pytest.log 2>&1
tail -4 gpurun_out/r2c7_pytest.log
run() {
  local name=$1; shift
  ( time timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c7_$name.json 2> gpurun_out/r2c7_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c7_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'phases', d['details'].get('phases_run_back_to_back_ms'), 'GBps', d['details'].get('halo_GBps_per_rank'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), d.get('lanczos',{}).get('lowest_ritz'), 'chk', d['details']['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c7_$name.err
}
BARGS="--no-e2e --exchange nccl" run nccl NCCL_DEBUG=WARN
BARGS="--no-extras --no-e2e --exchange pull" run pull X=1
