#!/bin/bash
# round 2, GPU call 26 (1 GPU): w0 = defaults (5 mover warps); w1 = k_divdamp grid sized by the occupancy query, k_acoustic_gather capped at 48 registers, evict-first loads / stores for the read-once strips of k_dt_edge; w2 = k_divdamp 48 registers, k_acoustic_gather 40, evict-first in k_dt_cellC
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in w0 w1 w2; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c26_$v.json 2> gpurun_out/c26_$v.err
done
python - <<P
import json
names=("k_divdamp","k_acoustic_gather","k_dt_edge","k_dt_cellC<false>","k_dt_cellC<true>","k_acoustic_lane<false>","k_acoustic_lane<true>")
print("variant step", *names)
for t in ("w0","w1","w2"):
    try:
        d=json.loads(open(f"gpurun_out/c26_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
