"""Summarise `ncu --set full` reports: per captured launch, the metrics the roofline discussion needs."""
import csv
import io
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__t_bytes.sum": "l1_bytes",
}


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return v * mult if mult else v


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print(path, "EMPTY")
        continue
    header, units = rows[0], rows[1]
    print(f"== {path}")
    for r in rows[2:]:
        d = dict(zip(header, r))
        u = dict(zip(header, units))
        out = {"kernel": d.get("Kernel Name", "")[:60]}
        for k, name in KEYS.items():
            if k in d and d[k] != "":
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if name in ("dram_read", "dram_write", "l2_bytes", "l1_bytes"):
                    v = to_bytes(v, u[k])
                if name == "duration":
                    v = to_us(v, u[k])
                out[name] = v
        dur = out.get("duration", 0)
        tr = out.get("dram_read", 0) + out.get("dram_write", 0)
        extra = f" traffic {tr / 1e6:.2f} MB -> {tr / dur / 1e3:.0f} GB/s" if dur else ""
        print("  ", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in out.items()}, extra)
