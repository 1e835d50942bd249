"""Launch list of one RK3 step from an `ncu --metrics gpu__time_duration.sum --csv` log, beside bench.py's CUDA-event figures:
    python profiles/launch_list.py gpurun_out/c27_launches.csv gpurun_out/c27_plain.log > profiles/r2_launch_list_final.md
The timed step = from the LAST k_setup_cell of the log to the end of that step (58 launches)."""
import csv, json, re, sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
def short(n):
    n = n.replace("void ", "")
    m = re.match(r"(\w+)(<[^>]*>)?", n)
    base, tpl = m.group(1), m.group(2) or ""
    tpl = re.sub(r", ?(\(int\))?0>", ">", tpl)
    tpl = tpl.replace("(bool)1", "true").replace("(bool)0", "false").replace("<1>", "<true>").replace("<0>", "<false>")
    return base + tpl
L = [(short(r[iN]), float(r[iV].replace(",", "")) * 1e-6) for r in data]          # ms
starts = [i for i, (n, _) in enumerate(L) if n == "k_setup_cell"]
s = starts[-1]
e = next((i for i in range(s + 1, len(L)) if L[i][0].startswith("k_summarize")), len(L))
step = L[s:e]
bench = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
ev = bench["kernels_ms_per_step"]
agg = OrderedDict()
for n, ms in step:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values()); tot_ev = sum(ev.values())
print(f"# Round 2, final library - ncu launch list of one RK3 step ({bench['config']['workload'].split(',')[0]} x 55 levels, 1 B200)\n")
print("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e`")
print(f"(the same command exited 0 without ncu immediately before: {bench['ms_per_step']:.2f} ms/step by CUDA events; `profiles/r2_call27.sh`).  Raw list: `profiles/r2_launches_final.csv`.")
print("Times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event figures of bench.py (`kernels_ms_per_step`), not absolutes.\n")
print(f"The timed step (last `k_setup_cell` of the log ... the check scan): {len(step)} launches, {tot:.2f} ms under ncu.\n")
print("| kernel | launches | ms (ncu) | share (ncu) | ms/step (CUDA events) | share (CUDA events) |\n|---|---|---|---|---|---|")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    e_ = ev.get(n)
    print(f"| {n} | {c} | {ms:.3f} | {ms / tot:.3f} | " + (f"{e_:.3f} | {e_ / tot_ev:.3f} |" if e_ is not None else "- | - |"))
