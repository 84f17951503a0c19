import json,sys
d=json.load(open(sys.argv[1]))
print("value", round(d["value"]/1e9,2), "G rows/s", round(d["ms_per_step"],2), "ms", d["step_ms"])
if d.get("e2e"): print("e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],2), "pinned", round(d["e2e"]["pinned_images"]["value"]/1e9,2))
for k in d["roofline"]["kernels"]: print("   ", k["kernel"], round(k["device_ms_per_step"],3), k["launches_per_step"], k.get("frac_of_peak"))
print("cpu", d.get("cpu_baseline"))
for w,v in (d.get("workloads") or {}).items():
    print(w, round(v["value"]/1e9,3), round(v["ms_per_step"],2), v.get("result_ok"), (v.get("e2e") or {}).get("value"), v.get("cpu_baseline"))
    for leg in ("sort","shuffle"):
        if leg in v: print("  ", leg, round(v[leg]["value"]/1e9,3), round(v[leg]["ms_per_step"],2), {k:v[leg][k] for k in v[leg] if k in ("exchange_ok","alltoall_gbs_per_rank","file_gbs","result_ok")}, [(k["kernel"], round(k["device_ms_per_step"],2), round(k.get("frac_of_peak") or 0,3)) for k in v[leg]["roofline"]["kernels"][:7]])
    if "roofline" in v: print("  ", [(k["kernel"], round(k["device_ms_per_step"],2), round(k.get("frac_of_peak") or 0,3)) for k in v["roofline"]["kernels"][:6]])
