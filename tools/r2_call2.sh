#!/bin/bash
# round 2, GPU call 2 (two B200): GPU test suite incl. the NCCL contexts, N=2 bench through the in-library context, exchange variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2c2_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2c2_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2c2_pytest.log 2>&1
tail -4 gpurun_out/r2c2_pytest.log
run() {  # name, extra env/args...
  local name=$1; shift
  ( time timeout 900 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c2_$name.json 2> gpurun_out/r2c2_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c2_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), 'halo', d['details'].get('halo_rows_max'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), 'ritz', d.get('lanczos',{}).get('lowest_ritz'), 'chk', d['details']['checksum_x_dot_Hx'])
    t=d.get('tri6x6')
    if t: print('   tri6x6 free', round(t['matrix_free']['ms_per_matvec'],2), 'csr', round(t['cached_csr']['ms_per_matvec'],3), t['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c2_$name.err
}
BARGS="" run full NCCL_DEBUG=WARN
BARGS="--no-extras --no-e2e" run allgather X=1 ; true
BARGS="--no-extras --no-e2e --exchange allgather" run allgather X=1
BARGS="--no-extras --no-e2e" run wrapranges EDCUDA_SHARD_POLICY=2
BARGS="--no-extras --no-e2e --chunks 4" run chunks4 X=1
BARGS="--no-extras --no-e2e --chunks 16" run chunks16 X=1
BARGS="--no-extras --no-e2e" run pull1 EDCUDA_PULL_STREAMS=1
BARGS="--no-extras --no-e2e" run pull4 EDCUDA_PULL_STREAMS=4
