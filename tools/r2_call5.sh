#!/bin/bash
# round 2, GPU call 5 (two B200): push exchange validation (tests, N=2 bench push vs pull), ncu captures on GPU 0
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2c5_pytest.log 2>&1
tail -5 gpurun_out/r2c5_pytest.log
run() {
  local name=$1; shift
  ( time timeout 900 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 10 --warmup 3 $BARGS ) > gpurun_out/r2c5_$name.json 2> gpurun_out/r2c5_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c5_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), 'phases', d['details'].get('phases_run_back_to_back_ms'), 'GBps', d['details'].get('halo_GBps_per_rank'), 'lanczos', d.get('lanczos',{}).get('ms_per_step'), d.get('lanczos',{}).get('lowest_ritz'), 'chk', d['details']['checksum_x_dot_Hx'])
except Exception as e:
    print('$name FAILED', e)
PY
  tail -2 gpurun_out/r2c5_$name.err
}
BARGS="--no-e2e --lanczos 100" run push X=1
BARGS="--no-extras --no-e2e --exchange pull" run pull X=1
BARGS="--no-extras --no-e2e --chunks 4" run push_c4 X=1
BARGS="--no-extras --no-e2e --chunks 16" run push_c16 X=1
# ---- ncu (one GPU): launch list of the headline bench command, full captures of the top kernels
export CUDA_VISIBLE_DEVICES=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_apply_u1 -s 3 -c 1 -f -o gpurun_out/r02_u1 python bench.py --steps 2 --no-extras --no-e2e > gpurun_out/r02_ncu_u1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k6b_canonicalize_nk -s 2 -c 1 -f -o gpurun_out/r02_k6nk python bench.py --workload tri6x6_k0A1_sz0 --steps 1 > gpurun_out/r02_ncu_k6nk.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_csr_blk -s 12 -c 4 -f -o gpurun_out/r02_spmv python tools/spmv_sweep.py 5300000 --lanes=4 > gpurun_out/r02_ncu_spmv.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k6c_combine -s 1 -c 1 -f -o gpurun_out/r02_k6c python bench.py --workload tri6x6_k0A1_sz0 --steps 1 > gpurun_out/r02_ncu_k6c.log 2>&1
ls -la gpurun_out/*.ncu-rep
