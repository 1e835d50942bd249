#!/bin/bash
# round 2, GPU call 9 (1 GPU): ablations of k_dt_edge_tile in the laboratory build -- consumer side alone (no copies, no wait) and staging alone
set -x
cd "$GRAFT_REPO_ROOT"
export MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_lab.so
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e --gather-stage 0"
for t in 1008 2008 1016 2016; do
  timeout 300 $B --edge-tiles $t > gpurun_out/c9_et$t.json 2> gpurun_out/c9_et$t.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c9_et$t.json").read().strip().splitlines()[-1])
k=d["kernels_ms_per_step"]
print("edge_tiles", $t, "step", d["ms_per_step"], {n:v for n,v in k.items() if "dt_edge" in n})
P
done
