#!/usr/bin/env bash
# Device-arm sweep of the pipelined DMR step over K1 tile sizes (DH_RRC_R) and K2 register caps (DH_DEMOD_MINB).
# usage: tools/sweep_step.sh > gpurun_out/sweep.txt
for R in 15 19 21 23 25; do
  for MB in 0 10 12 16; do
    line=$(DH_RRC_R=$R DH_DEMOD_MINB=$MB python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1)
    python - "$R" "$MB" <<PY
import json,sys
d=json.loads('''$line''')
r=d["roofline"]
print("R=%s MINB=%s  step %.4f ms  value %.1f  k1_in %.4f  k1_alone %.4f  k2_in %.3f k3_in %.3f" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"]/1e3, r["k1_ms_per_step"], r["not_overlapped"]["ms_per_launch"], r["stage_ms_per_step"]["k2_demod"], r["stage_ms_per_step"]["k3_k4_dmr"]))
PY
  done
done
