"""Summarise an `ncu --set full` report (no GPU needed):  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > table.md
One row per distinct kernel (mean over its captured launches)."""
import csv, io, subprocess, sys
from collections import OrderedDict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
M = OrderedDict([
    ("time us", "gpu__time_duration.sum"),
    ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("LTS %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("warps/SM", "sm__warps_active.avg.per_cycle_active"), ("regs", "launch__registers_per_thread"),
    ("fp64 %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("lg_thr", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("mio_thr", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
])
units = rows[1]
agg = OrderedDict()
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = []
    for k, m in M.items():
        if m not in ix:
            vals.append(float("nan")); continue
        v = float(r[ix[m]].replace(",", "")) if r[ix[m]] not in ("", "n/a") else float("nan")
        u = units[ix[m]]
        if k.startswith("DRAM") and k.endswith("MB"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if k == "time us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
        vals.append(v)
    agg.setdefault(name, []).append(vals)
print("| kernel | n | " + " | ".join(M) + " |")
print("|---|---|" + "---|" * len(M))
for name, vs in agg.items():
    mean = [sum(c) / len(c) for c in zip(*vs)]
    print(f"| `{name}` | {len(vs)} | " + " | ".join(f"{v:.1f}" if abs(v) >= 10 else f"{v:.2f}" for v in mean) + " |")
