set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_demod_split_gpu.py -x -q > gpurun_out/r03_split_tests.log 2>&1
tail -5 gpurun_out/r03_split_tests.log
timeout 300 tools/split_ab.sh "X=0" "DH_DEMOD_SPLIT=1" "DH_DEMOD_SPLIT=1 DH_DEMOD_PREFETCH=0" "DH_DEMOD_SPLIT=1 DH_RRC_R=19" "DH_DEMOD_SPLIT=1 DH_RRC_R=25" > gpurun_out/r03_split_ab.txt 2>&1
cat gpurun_out/r03_split_ab.txt
DH_DEMOD_SPLIT=1 timeout 200 python tools/pipe_throughput.py 20 > gpurun_out/r03_pipe_throughput_split.txt 2>&1
cat gpurun_out/r03_pipe_throughput_split.txt
