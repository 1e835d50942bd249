#!/bin/bash
# round 2, GPU call 22 (1 GPU): r0 = shipped; r1 = static rows staged in shared memory by lane-distributed loads (k_dt_edge Coriolis row,
# k_dt_theta_flux advection row), masks / row length requested together (k_acoustic_gather, k_smlstep + L2 prefetch of its own strips); r2 = k_acoustic_gather with its five rows in shared memory
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in r0 r1 r2; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c22_$v.json 2> gpurun_out/c22_$v.err
done
python - <<P
import json
names=("k_dt_edge","k_acoustic_gather","k_dt_theta_flux","k_smlstep","k_dt_cellC<false>")
print("variant step", *names)
for t in ("r0","r1","r2"):
    try:
        d=json.loads(open(f"gpurun_out/c22_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
