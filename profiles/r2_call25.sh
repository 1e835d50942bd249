#!/bin/bash
# round 2, GPU call 25 (1 GPU): k_acoustic_lane ablation -- u1: the sweeper skips its arithmetic (wrong results, measurement only): what is left is the data path; u2: 6 mover warps
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in u1 u2; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c25_$v.json 2> gpurun_out/c25_$v.err
done
python - <<P
import json
names=("k_acoustic_lane<false>","k_acoustic_lane<true>","k_dt_theta_flux","k_smlstep","k_dt_cellC<false>","k_dt_cellC<true>")
print("variant step", *names)
for t in ("u1","u2"):
    try:
        d=json.loads(open(f"gpurun_out/c25_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
