mkdir -p gpurun_out
( time timeout 280 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" ) > gpurun_out/c5_pytest_multi.log 2>&1; tail -40 gpurun_out/c5_pytest_multi.log
