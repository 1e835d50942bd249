#!/bin/bash
# round 2, GPU call 18 (1 GPU): stall-guided forms of the gather kernels (profiles/r2_stall_profiles.md) -- hoisted level-L / row-length
# loads, gathers of B slots requested together, useless advection-coefficient loads left out -- as variant libraries against the
# shipped one, x1.163842 x 55; every run must reproduce the checksum 68a7b531b35b42f9
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in old n0 n1 n2 n3 n4 n5 n6; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c18_$v.json 2> gpurun_out/c18_$v.err
done
python - <<P
import json
names=("k_dt_edge","k_dt_cellC<false>","k_dt_cellC<true>","k_acoustic_gather","k_dt_theta_flux","k_diag_cell","k_dt_cellA")
print("variant step", *names)
for t in ("old","n0","n1","n2","n3","n4","n5","n6"):
    try:
        d=json.loads(open(f"gpurun_out/c18_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
