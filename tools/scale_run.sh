# usage: bash tools/scale_run.sh N  -> full bench lines (L=32 with Lanczos, 6x6 triangular) on N GPUs into gpurun_out/
N=$1
run() { if [ "$N" = 1 ]; then python bench.py "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py "$@"; fi; }
run --gpus $N --steps 10 --warmup 3 --lanczos 50 2> gpurun_out/scale_l32_n$N.err | tail -1 > gpurun_out/scale_l32_n$N.json
run --gpus $N --workload tri6x6_k0A1_sz0 --steps 2 --warmup 1 2> gpurun_out/scale_tri_n$N.err | tail -1 > gpurun_out/scale_tri_n$N.json
python - <<PY
import json
for w in ("l32","tri"):
    try:
        d=json.load(open("gpurun_out/scale_%s_n$N.json" % w))
        print(w, "N=$N ms/step %.3f" % d["ms_per_step"], "kernel_ms", d["roofline"].get("kernel_ms"), "lanczos", d.get("lanczos",{}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "csr", d["config"].get("cached_csr",{}).get("ms_per_matvec"))
    except Exception as e:
        print(w, "failed", e); print(open("gpurun_out/scale_%s_n$N.err" % w).read()[-1500:])
PY
