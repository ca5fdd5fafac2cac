"""CPU: the JSON line of bench.py's reference arm carries the keys the driver's contract names (run on the smallest model)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny.en", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-tokens", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "cpu_baseline", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and line["value"] > 0
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["value"] == line["value"]
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_line_of_our_arm_meets_the_contract():
    """The GPU arm cannot run here; the line the last GPU run produced (profiles/r2_bench_default.json, written by bench.py itself) must carry
    every key of the driver's contract, be internally consistent, and cite an ncu traffic capture of the kernel source that is in the tree."""
    with open(os.path.join(ROOT, "profiles", "r2_bench_default.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        assert line["metric"] == json.load(f)["metric"].split(" @")[0]
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["warmup"] >= 3 and line["n_gpus"] == 1 and line["gpu_launches"] > 0 and "workload" in line["config"] and "l2" in line["config"]
    # value = audio seconds of one step / step time (16 chunks of 30 s)
    assert abs(line["value"] - 16 * 30.0 / (line["ms_per_step"] / 1e3)) < 1e-6 * line["value"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 16 * 480000 * 4 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= line["value"] * 1.001
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["alg_bytes_per_step"] / (r["ms_per_decode_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is not None and 0.9 < r["traffic"] / r["alg_bytes_per_step"] < 1.2
    with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
        cap = json.load(f)["bstep_kernel"]
    from faster_whisper_b200.build import source_fingerprint

    assert source_fingerprint(os.path.join(ROOT, "faster_whisper_b200", "csrc", cap["source_file"])) == cap["source_sha16"], \
        "the ncu capture is older than the kernel code: re-run tools/gpu_profile.sh"
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
