#!/usr/bin/env bash
# Round-2 single-GPU measurement batch (writes into gpurun_out/).
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_line_reference_arm.json 2>/dev/null
python bench.py --steps 4000 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_line_sustained.json 2>/dev/null
python bench.py --channels 8192 --steps 20 --warmup 5 > gpurun_out/r02_bench_line_8192ch.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rrc_fir|demod_kernel|dmr_kernel|pack_results" \
    --launch-skip 4 -c 4 -o gpurun_out/r02_full -f python tools/ncu_driver.py 4096 48000 3 dmr shard > gpurun_out/r02_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rrc_fir" --launch-skip 1 -c 1 -o gpurun_out/r02_full_f32 -f \
    python tools/ncu_driver.py 4096 48000 3 dmr f32 > gpurun_out/r02_full_f32.log 2>&1
ls -la gpurun_out | tail -12
