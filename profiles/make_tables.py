"""Render the per-kernel roofline table of DESIGN.md 4.3 from a bench.py JSON line:
    python profiles/make_tables.py profiles/r2_bench/<headline>.json
(algorithmic bytes from mpas_regent_b200/traffic.py; peak = MEASURED_PEAKS.json hbm_gbs, else the profiling guide's fallback)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpas_regent_b200 import traffic

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; src = "measured"
except Exception:
    peak = 6650.0; src = "fallback"
cl = 36044910 if "655362" in d["config"]["workload"] else None
assert cl, "headline mesh expected"
ms = d["kernels_ms_per_step"]; nl = d.get("kernels_launches_per_step", {})
seq = traffic.step_launches(True)
count = {k: seq.count(k) for k in set(seq)}
print(f"| kernel | launches / step | ms / step | ms / launch | units moved (contract + scratch) | GB/s on contract bytes | fraction of {peak:.0f} GB/s ({src}) |")
print("|---|---|---|---|---|---|---|")
tot = 0.0
for k, v in sorted(ms.items(), key=lambda kv: -kv[1]):
    key = traffic.lookup(k.strip("()"))
    n = nl.get(k) or count.get(key) or 0
    tot += v
    if key is None or not n:
        print(f"| `{k}` | {n or ''} | {v:.3f} | | | | |"); continue
    u_c, u_s = traffic.units(key, scratch=False), traffic.units(key, scratch=True)
    gbs = u_c * 8.0 * cl * n / (v * 1e-3) / 1e9
    print(f"| `{k}` | {n:g} | {v:.3f} | {v / n:.3f} | {u_c:g}" + (f" + {u_s - u_c:g}" if u_s != u_c else "") + f" | {gbs:.0f} | {gbs / peak:.2f} |")
print(f"| sum of kernels | | {tot:.2f} | | | | |")
print(f"| **step (CUDA events, no per-launch events)** | {d['gpu_launches'] / d['steps']:g} | **{d['ms_per_step']:.2f}** | | {traffic.SURVEY_STEP_UNITS_CANONICAL} (contract) | {d['step_hbm']['achieved_gbs_per_gpu']:.0f} | **{d['step_hbm']['frac_of_peak']:.2f}** |")
