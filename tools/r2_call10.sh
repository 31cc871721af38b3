#!/bin/bash
# round 2, GPU call 10 (two B200): side-stream pack + fence for the pull transport: correctness and N=2 timing, on / off
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --durations=5 ) > gpurun_out/r2c10_pytest.log 2>&1
tail -4 gpurun_out/r2c10_pytest.log
for side in 1 0; do
  ( time EDCUDA_SHARD_SIDE=$side timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$side bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2c10_n2_side$side.json 2> gpurun_out/r2c10_n2_side$side.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2c10_n2_side$side.json'))
print('side=$side ms', d['ms_per_step'], 'lanczos', (d.get('lanczos') or {}).get('ms_per_step'), (d.get('lanczos') or {}).get('lowest_ritz'), d['details'].get('phases_run_back_to_back_ms'), 'chk', d['details'].get('checksum'), d['clocks'])
PY
done
