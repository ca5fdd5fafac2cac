import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "stages_ms", "clocks")})
print("e2e", d.get("e2e"))
print("roofline", d.get("roofline"))
print("encoder", d.get("roofline_encoder"))
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
