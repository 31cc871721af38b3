# One-GPU check: column-blocked cached SpMV -- parity (forced blocking on small matrices, 6x6 full size) and timing
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cached_matrix or triangular_6x6" ) > gpurun_out/c8_pytest.log 2>&1; tail -4 gpurun_out/c8_pytest.log
B="--workload tri6x6_k0A1_sz0 --steps 2 --warmup 1"
timeout 300 python bench.py $B > gpurun_out/c8_tri_default.json 2> gpurun_out/c8_tri_default.err
EDCUDA_CSR_NOBLOCK=1 timeout 300 python bench.py $B > gpurun_out/c8_tri_noblock.json 2> gpurun_out/c8_tri_noblock.err
EDCUDA_CSR_BLOCK_COLS=2650000 timeout 300 python bench.py $B > gpurun_out/c8_tri_b8.json 2> gpurun_out/c8_tri_b8.err
EDCUDA_CSR_BLOCK_COLS=5300000 timeout 300 python bench.py $B > gpurun_out/c8_tri_b4.json 2> gpurun_out/c8_tri_b4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c8_tri_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        c = d["config"]["cached_csr"]
        print(f, "matrix-free ms %.1f | csr ms %.3f kernel_ms %.3f assemble %.2fs GB/s %.0f nnz %d" % (d["ms_per_step"], c["ms_per_matvec"], c["kernel_ms"], c["assemble_seconds"], c["spmv_GBps_per_gpu"], c["nnz"]))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-800:])
PY
