"""Experiment (round 1): which latency-hiding structure pays for a light gather kernel on B200.
Times the divergence-damping kernel variants (kernels.cuh, k_divdamp_v1..v5) on the headline mesh with
CUDA events, plus the per-kernel table of one full RK3 step.  Run on the GPU box:
    python profiles/exp_divdamp.py [cells] [levels]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mpas_regent_b200 import _abi, dynamics, traffic  # noqa: E402

nC = int(sys.argv[1]) if len(sys.argv) > 1 else 655362
L = int(sys.argv[2]) if len(sys.argv) > 2 else 55
mesh, st, t_init = bench.build_inputs(nC, L)
LIB = os.environ.get("MPASB200_LIB")          # optional: an alternative build of the library (launch-bound experiments)
g = dynamics.Dynamics(dynamics.dims_of(mesh, L), _abi.default_config(rkarg_policy=_abi.RKARG_STAGE_INDEX, acoustic_tma=int(os.environ.get("ACOUSTIC_TMA", "0"))), lib_path=LIB)
g.upload_mesh(st.static); g.upload_state(st.f, st.vert)
del st
dt = bench.dt_for(nC)
g.atm_compute_solve_diagnostics(False, -1)
for _ in range(2):
    g.atm_srk3(dt)
g.sync()
g.enable_kernel_timing(True); g.reset_kernel_timing()
for _ in range(1):
    g.atm_srk3(dt)
kt = g.kernel_times()
print("== per-kernel table, one RK3 step (ms per launch, GB/s algorithmic) ==")
tot = 0.0
for k, (ms, n) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    u = traffic.units(k, scratch=False) if k in traffic.K else float("nan")
    gbs = u * 8 * nC * L / (ms / n * 1e-3) / 1e9
    tot += ms / 1
    print(f"{k:22s} {ms / n:8.3f} ms x {n // 1:2d}/step = {ms / 1:7.3f} ms/step   {u:5.0f} units  {gbs:7.0f} GB/s")
print(f"total {tot:.2f} ms/step")
lib = g._lib
lib.mpasb200_debug_divdamp.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
lib.mpasb200_debug_divdamp.restype = C.c_int
bytes_dd = 9 * 8 * nC * L
lib = g._lib
if os.environ.get("ABLATE"):
    lib.mpasb200_debug_acoustic.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.mpasb200_debug_acoustic.restype = C.c_int
    print("== acoustic (TMA) ablations: 1 = no gathers, 2 = no sweep, 4 = no stores ==")
    for abl in (0, 1, 2, 3, 4, 7):
        g.reset_kernel_timing()
        for _ in range(6):
            assert lib.mpasb200_debug_acoustic(g._h, abl, 20.0) == 0
        (name, (ms, n)), = g.kernel_times().items()
        print(f"ablation {abl}: {ms / n:7.3f} ms")
if os.environ.get("SKIP_VARIANTS"):
    sys.exit(0)
print("== divdamp variants ==")
for var, arg, label in [(0, 0, "production"), (1, 0, "v1 flag+ecv together"), (2, 1332, "v2 +L2 prefetch 1332 ahead"),
                        (2, 2664, "v2 +L2 prefetch 2664 ahead"), (3, 0, "v3 two edges/thread"), (4, 148 * 9, "v4 persistent 1332 blocks"),
                        (4, 148 * 18, "v4 persistent 2664 blocks"), (4, 148 * 36, "v4 persistent 5328 blocks"), (5, 0, "v5 four levels/thread"),
                        (6, 333, "v6 L2 data prefetch 333 ahead"), (6, 666, "v6 L2 data prefetch 666 ahead"), (6, 1332, "v6 L2 data prefetch 1332 ahead"),
                        (6, 2664, "v6 L2 data prefetch 2664 ahead"), (7, 148 * 8, "v7 persistent + 4 levels, 1184 blocks"),
                        (7, 148 * 16, "v7 persistent + 4 levels, 2368 blocks"),
                        (8, 148 * 9, "v8 persistent, no prefetch, 1332 blocks"), (8, 148 * 18, "v8 persistent, no prefetch, 2664 blocks"),
                        (8, 148 * 72, "v8 persistent, no prefetch, 10656 blocks")]:
    g.reset_kernel_timing()
    for _ in range(12):
        rc = lib.mpasb200_debug_divdamp(g._h, var, 100.0, arg)
        assert rc == 0, rc
    t = g.kernel_times()
    (name, (ms, n)), = t.items()
    print(f"{label:32s} {ms / n:7.3f} ms  {bytes_dd / (ms / n * 1e-3) / 1e9:7.0f} GB/s")
