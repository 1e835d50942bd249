#!/bin/bash
# round 2, GPU call 24 (1 GPU): t0 = shipped; k_acoustic_lane with 5 mover warps (t1) and 2 (t2) instead of 3: is the pipeline bound by its movers or by its sweeper?
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in t0 t1 t2; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c24_$v.json 2> gpurun_out/c24_$v.err
done
python - <<P
import json
names=("k_acoustic_lane<false>","k_acoustic_lane<true>","k_dt_theta_flux","k_smlstep","k_dt_cellC<false>","k_dt_cellC<true>")
print("variant step", *names)
for t in ("t0","t1","t2"):
    try:
        d=json.loads(open(f"gpurun_out/c24_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
