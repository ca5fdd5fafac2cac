"""Turn the scratch output of a profiling gpurun (gpurun_out/) into the committed evidence under profiles/ (round 2).

    python tools/make_profiles.py            # after tools/gpu_profile.sh and tools/gpu_final.sh ran on the GPU box
"""
import csv
import gzip
import hashlib
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_whisper_b200.build import source_fingerprint  # noqa: E402
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = "r2"


def sha16(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return dict(zip(rows[0], zip(rows[2], rows[1])))


def top_lines(rep, n=30):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    hdr, fpath, items, tot = None, None, [], 0
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and fpath and r[0].isdigit() and int(r[0]) > 0:
            try:
                c = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            items.append((c, os.path.basename(fpath), int(r[0]), r[1][:110]))
            tot += c
    items.sort(reverse=True)
    return tot, items[:n]


def full_report(name, kernel_file):
    rep = os.path.join(OUT, f"full_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    d = raw_metrics(rep)

    def val(k):
        return d.get(k, ("", ""))

    lines = [f"ncu --set full --clock-control none --import-source on, one launch of {name}_kernel (large-v3 synthetic weights; tools/gpu_profile.sh)", ""]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.sum.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
    for k in keys:
        v, u = val(k)
        lines.append(f"{k:75s} {v} {u}")
    stalls = sorted(((int(float(v[0])), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v[0] not in ("", "n/a")), reverse=True)
    tot = sum(c for c, _ in stalls) or 1
    lines += ["", "warp stall samples (all warps, incl. the producer / MMA warps that spin on mbarriers by design):"]
    lines += [f"  {c:8d} {100 * c / tot:5.1f}%  {k}" for c, k in stalls[:12]]
    tsum, tl = top_lines(rep)
    lines += ["", f"hottest source lines by samples (of {tsum}):"]
    lines += [f"  {c:7d} {100 * c / max(tsum, 1):5.1f}%  {f}:{ln}: {s}" for c, f, ln, s in tl]
    with open(os.path.join(PROF, f"{TAG}_ncu_full_{name}.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    dur = float(val("gpu__time_duration.sum")[0].replace(",", ""))
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = float(rd[0].replace(",", "")) * mult.get(rd[1], 1) + float(wr[0].replace(",", "")) * mult.get(wr[1], 1)
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(val("gpu__time_duration.sum")[1], 1.0)
    return {"dram_bytes_per_launch": int(traffic), "duration_us": round(dur_us, 2),
            "source_file": kernel_file, "source_sha16": source_fingerprint(os.path.join(ROOT, "faster_whisper_b200", "csrc", kernel_file)),
            "capture": f"profiles/{TAG}_ncu_full_{name}.txt (ncu --set full, one launch, dram__bytes_read.sum + dram__bytes_write.sum)"}


def main():
    os.makedirs(PROF, exist_ok=True)
    traffic = {}
    for name, src in (("bstep", "bstep.cu"), ("dstep", "dstep.cu")):
        ent = full_report(name, src)
        if ent:
            traffic[f"{name}_kernel"] = ent
    if traffic:
        with open(os.path.join(PROF, f"{TAG}_ncu_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    lst = os.path.join(OUT, "launches_batched.csv")
    if os.path.exists(lst):
        with open(lst, "rb") as f, gzip.open(os.path.join(PROF, f"{TAG}_launches_batched.csv.gz"), "wb") as g:
            shutil.copyfileobj(f, g)
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), lst], capture_output=True, text=True).stdout
        with open(os.path.join(PROF, f"{TAG}_launches_batched_summary.txt"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary`\n"
                    "(the default batched line: 16 chunks per step; per-launch times are cold-cache and serialised — compare SHARES, not absolutes)\n\n" + summ)
    for log, out, kern in (("step_prof.log", "bstep_phase_profile", "bstep"), ("step_prof_single.log", "dstep_phase_profile", "dstep")):
        err = os.path.join(OUT, log)
        if os.path.exists(err):
            lines = [ln for ln in open(err) if f"{kern} prof" in ln or "decode" in ln]
            with open(os.path.join(PROF, f"{TAG}_{out}.txt"), "w") as f:
                f.write(f"B2W_LIBRARY=libb200whisper_ticks.so python tools/step_ab.py --prof (large-v3, prompt 4 + 128 tokens): device timers of {kern}_kernel, CTA 0\n"
                        "(last decode step, t = 132; %globaltimer at every grid barrier = work / barrier wait per phase; clock64 counters inside the phases —\n"
                        "the counters are compiled in only in the ticks build, which is ~3 % slower than the default one)\n\n" + "".join(lines[-22:]))
    for src, dst in (("bench_default.json", "bench_default"), ("bench_reference.json", "bench_reference"), ("bench_b16_int8.json", "bench_batched_int8"),
                     ("bench_b1_dstep.json", "bench_single")):
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(PROF, f"{TAG}_{dst}.json"))
    t = os.path.join(OUT, "test_all_gpu.log")
    if os.path.exists(t):
        shutil.copy(t, os.path.join(PROF, f"{TAG}_pytest_gpu.txt"))
    census = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_census.py")], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, f"{TAG}_sass_census.txt"), "w") as f:
        f.write(census)
    print("profiles written:", sorted(x for x in os.listdir(PROF) if x.startswith(TAG)))


if __name__ == "__main__":
    main()
