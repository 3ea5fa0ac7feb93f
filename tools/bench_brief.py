"""Prints the key numbers of bench.py JSON lines read from stdin (tuning aid)."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    st = r.get("stage_ms_per_step") or {}
    e = d.get("e2e") or {}
    print("value %.1f Ms/s  step %.4f ms  k1 %.4f k2 %.4f k3 %.4f  frac %.4f  e2e %s  pipelining=%s" % (
        d["value"], d["ms_per_step"], st.get("k1_rrc", 0), st.get("k2_demod", 0), st.get("k3_k4_dmr", 0),
        r.get("frac") or 0, ("%.1f" % e["value"]) if e else "-", (d.get("config") or {}).get("pipelining")))
