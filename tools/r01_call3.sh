# One-GPU check: byte-offset ELL table (default) vs EXACT per-slab slot counts (EDCUDA_U1_ELLX=1): parity subset + timing
mkdir -p gpurun_out
K="fast_path or apply_heisenberg or other_sectors or wrap_aware or segmented or full_size_properties or lanczos or falls_back"
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" ) > gpurun_out/c3_pytest_default.log 2>&1; tail -3 gpurun_out/c3_pytest_default.log
( EDCUDA_U1_ELLX=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" ) > gpurun_out/c3_pytest_ellx.log 2>&1; tail -3 gpurun_out/c3_pytest_ellx.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for i in 1 2; do
timeout 200 python bench.py $B > gpurun_out/c3_bench_base_$i.json 2> gpurun_out/c3_bench_base_$i.err
EDCUDA_U1_ELLX=1 timeout 200 python bench.py $B > gpurun_out/c3_bench_ellx_$i.json 2> gpurun_out/c3_bench_ellx_$i.err
done
EDCUDA_U1_ELLX=1 timeout 200 python bench.py $B --workload j1j2_chain_L28_sz0 > gpurun_out/c3_bench_ellx_j1j2.json 2>&1
timeout 200 python bench.py $B --workload j1j2_chain_L28_sz0 > gpurun_out/c3_bench_base_j1j2.json 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c3_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f kernel_ms %.3f frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]), "chk", d["config"].get("checksum_x_dot_Hx"))
    except Exception as e:
        print(f, "FAILED", e)
PY
