set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_shard_gpu.py tests/test_state_gpu.py tests/test_ysf_gpu.py -m gpu -q > gpurun_out/r03_pytest_gpu_rest.log 2>&1
tail -4 gpurun_out/r03_pytest_gpu_rest.log
