#!/bin/bash
# round 2, GPU call 3 (two B200): full GPU suite (NCCL contexts included), N=2 bench with interior-first chunks, config 4 kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2c3_pytest.log 2>&1
tail -6 gpurun_out/r2c3_pytest.log
( time EDCUDA_K6_TIMING=1 timeout 600 python bench.py --workload tri6x6_k0A1_sz0 --steps 3 ) > gpurun_out/r2c3_tri_n1.json 2> gpurun_out/r2c3_tri_n1.err
grep "K6 staged" gpurun_out/r2c3_tri_n1.err | tail -2
python -c "
import json; d=json.load(open('gpurun_out/r2c3_tri_n1.json'))['tri6x6']; print('tri N=1 free', d['matrix_free']['ms_per_matvec'], 'csr', d['cached_csr']['ms_per_matvec'], d['cached_csr']['assemble_seconds'], d['checksum_x_dot_Hx'])"
( time EDCUDA_K6_NONECKLACE=1 EDCUDA_CSR_NOCODE=1 EDCUDA_K6_TIMING=1 timeout 600 python bench.py --workload tri6x6_k0A1_sz0 --steps 3 ) > gpurun_out/r2c3_tri_n1_old.json 2> gpurun_out/r2c3_tri_n1_old.err
grep "K6 staged" gpurun_out/r2c3_tri_n1_old.err | tail -1
python -c "
import json; d=json.load(open('gpurun_out/r2c3_tri_n1_old.json'))['tri6x6']; print('tri N=1 (old canonicalize, uncoded) free', d['matrix_free']['ms_per_matvec'], 'csr', d['cached_csr']['ms_per_matvec'])"
run() {
  local name=$1; shift
  ( time timeout 900 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c3_$name.json 2> gpurun_out/r2c3_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c3_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), 'halo', d['details'].get('halo_rows_max'), 'phases', d['details'].get('phases_run_back_to_back_ms'), 'pullGBps', d['details'].get('pull_GBps_per_rank'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), 'chk', d['details']['checksum_x_dot_Hx'], 'clocks', d.get('clocks'))
    t=d.get('tri6x6')
    if t: print('   tri6x6 free', round(t['matrix_free']['ms_per_matvec'],2), 'csr', round(t['cached_csr']['ms_per_matvec'],3), t['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c3_$name.err
}
BARGS="" run full X=1
BARGS="--no-extras --no-e2e --chunks 4" run chunks4 X=1
BARGS="--no-extras --no-e2e --chunks 12" run chunks12 X=1
BARGS="--no-extras --no-e2e" run pull2 EDCUDA_PULL_STREAMS=2
