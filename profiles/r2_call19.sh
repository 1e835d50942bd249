#!/bin/bash
# round 2, GPU call 19 (1 GPU): second set of stall-guided forms -- slot 0 of the gather loops peeled to the top (k_acoustic_gather,
# k_dt_theta_flux, k_diag_cell, k_dt_edge), last-edge static for w_adv_curv, level-L hoist in k_vert_imp, L1 instead of L2 prefetch,
# cross-block row prefetch; p0 = the defaults adopted after call 18
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in p0 p1 p2 p3; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c19_$v.json 2> gpurun_out/c19_$v.err
done
python - <<P
import json
names=("k_dt_edge","k_dt_cellC<false>","k_dt_cellC<true>","k_acoustic_gather","k_dt_theta_flux","k_diag_cell","k_dt_cellA","k_vert_imp")
print("variant step", *names)
for t in ("p0","p1","p2","p3"):
    try:
        d=json.loads(open(f"gpurun_out/c19_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
