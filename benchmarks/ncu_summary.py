"""`.ncu-rep` (--set full) -> the per-launch summary CSV committed under profiles/:
    python benchmarks/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_ncu_step_kernels_x.csv
DRAM fractions are computed from bytes / duration (the `dram__throughput…pct` metric returns no data with this ncu on
B200); the copy peak is MEASURED_PEAKS.json's."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

rep, out = sys.argv[1], sys.argv[2]
peak = json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json")).get("hbm_gbs", 6457.4)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(head)}


def val(r, name, default=""):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return r[i]


def scaled(r, name, to):
    """value of a metric converted to `to` ('us' or 'MB') from whatever unit ncu printed"""
    v = val(r, name, None)
    if v is None or isinstance(v, str):
        return ""
    u = units[col[name]]
    f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    return v * f


fields = ["kernel", "grid", "block", "duration_us", "dram_read_MB", "dram_write_MB", "regs", "warps_active_pct", "sm_pct",
          "warp_inst_executed", "waves_per_sm", "l2_hit_pct", "stall_long_scoreboard", "dram_gbs_computed",
          f"dram_frac_of_measured_copy_peak_{int(peak)}", "dram_frac_of_nominal_8000"]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(fields)
    for r in body:
        dur = scaled(r, "gpu__time_duration.sum", "us")
        rd, wr = scaled(r, "dram__bytes_read.sum", "MB"), scaled(r, "dram__bytes_write.sum", "MB")
        gbs = (rd + wr) / dur * 1e3 if dur else 0.0
        w.writerow([r[col["Kernel Name"]], val(r, "launch__grid_size"), val(r, "launch__block_size"), round(dur, 3),
                    round(rd, 3), round(wr, 3), val(r, "launch__registers_per_thread"),
                    val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), val(r, "smsp__inst_executed.sum"),
                    val(r, "launch__waves_per_multiprocessor"), val(r, "lts__t_sector_hit_rate.pct"),
                    val(r, "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
                        val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")),
                    round(gbs, 1), round(gbs / peak, 3), round(gbs / 8000.0, 3)])
print("wrote", out, len(body), "launches")
