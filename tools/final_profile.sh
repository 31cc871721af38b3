# Round-end evidence run on one B200: tests, bench line (with Lanczos), ncu launch list of the bench, full ncu captures.
mkdir -p gpurun_out
set -x
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --lanczos 50 > gpurun_out/bench_r01_n1_final.json 2> gpurun_out/bench_r01_n1_final.err
tail -c 700 gpurun_out/bench_r01_n1_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference_arm.json 2> gpurun_out/bench_r01_reference_arm.err
tail -c 400 gpurun_out/bench_r01_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_launches_final_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_apply_u1 -s 3 -c 1 -f -o gpurun_out/u1_v9 python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_u1_v9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv_csr_blk -s 13 -c 1 -f -o gpurun_out/spmv_blk python tools/spmv_sweep.py 5300000 > gpurun_out/ncu_spmv_blk.log 2>&1
ls -la gpurun_out/*.ncu-rep
