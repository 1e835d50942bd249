#!/bin/bash
# round 2, GPU call 28 (1 GPU): k_dt_edge is now at 79 % of the L1 data pipe (prof_r2f): one WARP per column -- block (32, C), 4 idle lanes -- for fewer wavefronts per gathered column; x0 = shipped (28 x 8), x1 / x2 / x3 = C = 7 / 8 / 6
set -x
cd "$GRAFT_REPO_ROOT"
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for v in x0 x1 x2 x3; do
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_$v.so timeout 300 $B > gpurun_out/c28_$v.json 2> gpurun_out/c28_$v.err
done
python - <<P
import json
names=("k_dt_edge","k_dt_cellC<false>","k_acoustic_gather","k_divdamp")
print("variant step", *names)
for t in ("x0","x1","x2","x3"):
    try:
        d=json.loads(open(f"gpurun_out/c28_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
        print(t, round(d["ms_per_step"],3), *[k.get(n) for n in names], d["check"]["combined_checksum"])
    except Exception as e:
        print(t, "FAILED", e)
P
