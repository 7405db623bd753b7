import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f img/s  ms/step %.3f  roofline frac %.3f (%.0f GB/s)" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved"]))
if d.get("e2e"): print("e2e", d["e2e"].get("value"), " resident:", (d.get("e2e_resident_pred") or {}).get("value"))
print("clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
for k, v in (d.get("extra") or {}).items():
    print("  %-36s %9.1f us %7.0f GB/s  frac %.3f" % (k, v["us"], v["GBps"], v["frac_of_hbm_peak"]))
