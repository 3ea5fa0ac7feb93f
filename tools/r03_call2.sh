set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_demod_split_gpu.py -q > gpurun_out/r03_split_tests.log 2>&1
tail -5 gpurun_out/r03_split_tests.log
timeout 200 python tools/split_small_banks.py 20 > gpurun_out/r03_split_small_banks.txt 2>&1
cat gpurun_out/r03_split_small_banks.txt
DH_DEMOD_SPLIT=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r03_split_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r03_split_under_ncu.log 2>&1
grep -E "demod_|rrc_fir|dmr_kernel" gpurun_out/r03_split_launches.csv | tail -12
