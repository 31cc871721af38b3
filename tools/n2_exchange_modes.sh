# Two-GPU run: L=32 matvec + Lanczos with the two exchange modes (p2p peer loads / dma split exchange) and the all-gather baseline
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for ex in p2p dma allgather; do
  timeout 300 $T bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex --lanczos 30 > gpurun_out/c4_n2_$ex.json 2> gpurun_out/c4_n2_$ex.err
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" > gpurun_out/c4_pytest_multi.log 2>&1; tail -3 gpurun_out/c4_pytest_multi.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c4_n2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f kernel_ms %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]), "lanczos", (d.get("lanczos") or {}).get("ms_per_step"),
              "ritz", (d.get("lanczos") or {}).get("lowest_ritz"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "chk", d["config"].get("checksum_x_dot_Hx"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
