#!/usr/bin/env bash
# A/B of the K2 schedules (dh_demod_set_split) on the pipelined DMR step, device arm only.
# usage: tools/split_ab.sh > gpurun_out/split_ab.txt
run() {
  line=$(env "$@" python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1)
  python - "$*" <<PY
import json,sys
d=json.loads('''$line''')
r=d["roofline"]; s=r["stage_ms_per_step"]
i=r["not_overlapped"]["stage_ms_per_step"]
print("%-44s step %.4f ms  value %.1f  k1_in %.4f k2_in %.3f k3_in %.3f | alone k1 %.4f k2 %.3f k3 %.3f | launches %d" % (
    sys.argv[1], d["ms_per_step"], d["value"]/1e3, r["k1_ms_per_step"], s["k2_demod"], s["k3_k4_dmr"],
    i["k1_rrc"], i["k2_demod"], i["k3_k4_dmr"], d["gpu_launches"]))
PY
}
for cfg in "$@"; do run $cfg; done
