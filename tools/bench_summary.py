import sys, json
for line in sys.stdin:
    line = line.strip()
    if line.startswith("=="):
        print(line); continue
    if not line.startswith("{"):
        continue
    try:
        d = json.loads(line); r = d["roofline"]
        print("  rays/s %.0f  ms/step %.1f  e2e %.0f  roofline(pre) %.1f TF = %.3f  whole %.3f  stages %s launches %d clk %s %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["frac"], r["whole_step_frac"],
            {k: round(v, 1) for k, v in r["stage_ms_per_step"].items()}, d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
        if "cpu_baseline" in d: print("  cpu_baseline", d["cpu_baseline"])
    except Exception as e:
        print("  ??", e, line[:200])
