# Round-end evidence run on one B200: tests, bench line (with Lanczos), ncu launch list of the bench, one full ncu capture.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --lanczos 100 > gpurun_out/bench_r01_n1_final.json 2> gpurun_out/bench_r01_n1_final.err
tail -c 600 gpurun_out/bench_r01_n1_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_launches_final_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_apply_u1 -s 3 -c 1 -f -o gpurun_out/u1_v8 python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_u1_v8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k6b_canonicalize_tr -s 60 -c 1 -f -o gpurun_out/k6b_tr python bench.py --workload tri6x6_k0A1_sz0 --steps 1 --warmup 1 > gpurun_out/ncu_k6b_tr.log 2>&1
ls -la gpurun_out/*.ncu-rep
