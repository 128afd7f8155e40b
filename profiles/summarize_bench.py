import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value %.3e ms %.4f kernel_ms %.4f frac %.3f e2e %.3e launches %d idx %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"],d["e2e"]["value"],d["gpu_launches"], d.get("index")))
    for k,v in d.get("extras",{}).items(): print("   ",k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
