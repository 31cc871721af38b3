#!/bin/bash
# round 2, GPU call 1 (one B200): GPU test suite, smoke, default bench line, launch-order experiments of the u1 kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
nproc >> gpurun_out/r2c1_gpu.txt; free -g >> gpurun_out/r2c1_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2c1_pytest.log 2>&1
tail -5 gpurun_out/r2c1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1; tail -2 gpurun_out/r2c1_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -c 600 gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_bench.err
for ord in class key:1,-1,16 key:1,8,16 key:0,-1,17; do
  EDCUDA_U1_ORDER=$ord timeout 300 python bench.py --no-extras --no-e2e --steps 10 > gpurun_out/r2c1_order_$(echo $ord | tr ':,' '__').json 2>> gpurun_out/r2c1_order.err
  echo "$ord: $(python -c "import json,sys; d=json.load(open('gpurun_out/r2c1_order_$(echo $ord | tr ':,' '__').json')); print(d['ms_per_step'], d['details']['checksum_x_dot_Hx'])" 2>&1)"
done
timeout 300 python bench.py --no-extras --no-e2e --steps 10 --workload j1j2_chain_L28_sz0 > gpurun_out/r2c1_j1j2.json 2>> gpurun_out/r2c1_order.err
EDCUDA_U1_ORDER=class timeout 300 python bench.py --no-extras --no-e2e --steps 10 --workload j1j2_chain_L28_sz0 > gpurun_out/r2c1_j1j2_class.json 2>> gpurun_out/r2c1_order.err
python -c "
import json
for f in ('r2c1_j1j2','r2c1_j1j2_class'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'])"
