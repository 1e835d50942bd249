#!/bin/bash
# round 2, GPU call 20 (1 GPU): q0 = defaults after call 19; q1 = cross-block row prefetch in every gather kernel; q2 = in-thread L2
# prefetch of the columns the slot loops gather (lanes spread over slot x line); q3 = both, without the tail prefetch of k_dt_edge
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in q0 q1 q2 q3; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c20_$v.json 2> gpurun_out/c20_$v.err
done
python - <<P
import json
names=("k_dt_edge","k_dt_cellC<false>","k_dt_cellC<true>","k_acoustic_gather","k_dt_theta_flux","k_diag_cell","k_dt_cellA","k_dt_cellB","k_dt_edge_euler","k_diag_edge<false>","k_diag_edge<true>")
print("variant step", *names)
for t in ("q0","q1","q2","q3"):
    try:
        d=json.loads(open(f"gpurun_out/c20_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
