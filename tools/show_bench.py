import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    print("==", path)
    print({k: d.get(k) for k in ("impl", "value", "ms_per_step", "gpu_launches", "stages_ms", "device_ms_per_step", "clocks")})
    print("e2e", d.get("e2e"))
    print("roofline", d.get("roofline"))
    print("encoder", d.get("roofline_encoder"))
    if "cpu_baseline" in d:
        print("cpu", d["cpu_baseline"])
