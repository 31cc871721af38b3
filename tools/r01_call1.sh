# One-GPU check run: full GPU test suite, default bench line, then the TMA-staging knob (parity subset + timing).
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1; tail -5 gpurun_out/c1_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --lanczos 50 > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err; tail -c 1500 gpurun_out/c1_bench_default.json
( EDCUDA_U1_TMA=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or apply_heisenberg or other_sectors or wrap_aware or segmented or full_size_properties or lanczos" ) > gpurun_out/c1_pytest_tma.log 2>&1; tail -5 gpurun_out/c1_pytest_tma.log
for i in 1 2; do
EDCUDA_U1_TMA=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_bench_tma_$i.json 2> gpurun_out/c1_bench_tma_$i.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_bench_base_$i.json 2> gpurun_out/c1_bench_base_$i.err
done
EDCUDA_U1_TMA=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --lanczos 50 > gpurun_out/c1_bench_tma_lanczos.json 2> gpurun_out/c1_bench_tma_lanczos.err
EDCUDA_U1_TMA=1 timeout 200 python bench.py --workload j1j2_chain_L28_sz0 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_bench_tma_j1j2.json 2>&1
timeout 200 python bench.py --workload j1j2_chain_L28_sz0 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_bench_base_j1j2.json 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c1_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f kernel_ms %.3f frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]),
              "lanczos", (d.get("lanczos") or {}).get("ms_per_step"), "ritz", (d.get("lanczos") or {}).get("lowest_ritz"), "chk", d["config"].get("checksum_x_dot_Hx"))
    except Exception as e:
        print(f, "FAILED", e)
PY
