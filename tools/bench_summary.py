import sys, json
for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        if ln: print(ln)
        continue
    d = json.loads(ln)
    r = d["roofline"]; c = d.get("clocks") or {}
    print(d["config"]["config"], d["run"]["kernel"], "us/launch %.3f" % (r["avg_launch_ms"] * 1e3), "GB/s %.0f" % r["achieved"], "frac %.3f" % r["frac"],
          "Mpix/s %.0f" % d["value"], "clk", c.get("sm_mhz"), "W", c.get("power_w_max"), c.get("reasons"), "passes", d["run"]["passes_per_step"], "ok", d["parity_spot_check"])
