set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r03_pytest_gpu.log 2>&1
tail -4 gpurun_out/r03_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03_smoke.log 2>&1
tail -2 gpurun_out/r03_smoke.log
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool python tools/sanitize_split.py > gpurun_out/r03_sanitize_$tool.log 2>&1
  grep -E "SUMMARY|done" gpurun_out/r03_sanitize_$tool.log
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r03_bench_line.json 2> gpurun_out/r03_bench_line.err
tail -c 600 gpurun_out/r03_bench_line.json
