"""Summarise a launch timeline written by `bench.py --timeline FILE` (mpasb200_enable_kernel_timing(h, 2)):
    python profiles/timeline_summary.py gpurun_out/timeline_n2.json > profiles/r2_timeline_2gpu.md
Streams: 0 = compute, 1 = the handle's communication stream (k_pack, nccl_send_recv, k_unpack and the sent-cell launches)."""
import json, sys

d = json.load(open(sys.argv[1]))
E = sorted(d["entries"], key=lambda e: e[1])
t_first, t_last = min(e[1] for e in E), max(e[2] for e in E)


def union(iv):
    iv = sorted(iv); out = []
    for a, b in iv:
        if out and a <= out[-1][1]: out[-1][1] = max(out[-1][1], b)
        else: out.append([a, b])
    return out


def length(iv): return sum(b - a for a, b in iv)


def intersect(x, y):
    out = []; i = j = 0
    while i < len(x) and j < len(y):
        a, b = max(x[i][0], y[j][0]), min(x[i][1], y[j][1])
        if a < b: out.append([a, b])
        if x[i][1] < y[j][1]: i += 1
        else: j += 1
    return out


comp = union([(e[1], e[2]) for e in E if e[3] == 0])
comm = union([(e[1], e[2]) for e in E if e[3] == 1])
xchg = union([(e[1], e[2]) for e in E if e[3] == 1 and e[0] in ("k_pack", "k_unpack", "nccl_send_recv")])
both = intersect(comp, comm)
hidden = intersect(xchg, comp)
print(f"# Round 2 — launch timeline of one RK3 step on {d['n_gpus']} GPUs (x1.{d['mesh']} × {d['levels']} levels, rank 0)\n")
print("`nsys` is not in this image; the timeline is taken by the library itself: `mpasb200_enable_kernel_timing(h, 2)` brackets every launch and")
print("every NCCL send/recv group with a CUDA-event pair on the stream it runs on and reports start / end against one origin")
print("(`bench.py --timeline`, `profiles/timeline_summary.py`).  The event pairs serialise nothing but add ~2 µs per launch.\n")
print(f"* span of the step: {t_last - t_first:.3f} ms; {len(E)} timed intervals ({sum(1 for e in E if e[3] == 1)} on the communication stream)")
print(f"* compute stream busy {length(comp):.3f} ms; communication stream busy {length(comm):.3f} ms, of which concurrent with compute {length(both):.3f} ms")
print(f"* exchanges proper (`k_pack` + `nccl_send_recv` + `k_unpack`): {length(xchg):.3f} ms in {sum(1 for e in E if e[0] == 'nccl_send_recv')} groups; **{length(hidden):.3f} ms ({100 * length(hidden) / max(length(xchg), 1e-9):.0f} %) under compute-stream kernels**, {length(xchg) - length(hidden):.3f} ms exposed")
gaps = [(comp[i + 1][0] - comp[i][1]) for i in range(len(comp) - 1)]
print(f"* compute-stream gaps: {len(gaps)} totalling {sum(gaps):.3f} ms (largest {max(gaps) if gaps else 0:.3f} ms)\n")
# the first acoustic loop iteration with an exchange, as a table
idx = [i for i, e in enumerate(E) if e[0] == "nccl_send_recv"]
if idx:
    k = idx[min(2, len(idx) - 1)]
    t0 = E[k][1] - 0.25; t1 = E[k][2] + 0.35
    print(f"One acoustic-loop exchange in detail (everything that starts between {t0 - t_first:.3f} and {t1 - t_first:.3f} ms of the step):\n")
    print("| start ms | end ms | stream | launch |\n|---|---|---|---|")
    for e in E:
        if t0 <= e[1] <= t1:
            print(f"| {e[1] - t_first:.3f} | {e[2] - t_first:.3f} | {'comm' if e[3] else 'compute'} | `{e[0]}` |")
