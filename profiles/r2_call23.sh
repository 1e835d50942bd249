#!/bin/bash
# round 2, GPU call 23 (1 GPU): s0 = defaults (k_dt_theta_flux with its advection row in shared memory); s1 = k_dt_cellC theta-loop rows in shared
# memory, theta_flux unroll 3, k_smlstep masks requested together; s2 = theta_flux unroll 5, k_smlstep + prefetch of its used slots; s3 = theta_flux unroll 1, cellC rows + 5 blocks
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in s0 s1 s2 s3; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c23_$v.json 2> gpurun_out/c23_$v.err
done
python - <<P
import json
names=("k_dt_theta_flux","k_smlstep","k_dt_cellC<false>","k_dt_cellC<true>","k_dt_edge")
print("variant step", *names)
for t in ("s0","s1","s2","s3"):
    try:
        d=json.loads(open(f"gpurun_out/c23_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
