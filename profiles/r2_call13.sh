#!/bin/bash
# round 2, GPU call 13 (1 GPU): init_atm_case_jw test; own-cell reuse in the edgesOnCell loops (shipped: k_acoustic_gather only;
# v0 = none; v1 = also k_dt_cellA/B/C + the L2 prefetch in k_dt_edge) -- bit identity through the checksum, per-kernel times
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "init_atm_case_jw or acoustic_modes or task_parity" > gpurun_out/c13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest.log
grep -E "worst relative|passed|failed|rc=" gpurun_out/c13_pytest.log | tail -8
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
run() {
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/$2 timeout 300 $B > gpurun_out/c13_$1.json 2> gpurun_out/c13_$1.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c13_$1.json").read().strip().splitlines()[-1])
k=d["kernels_ms_per_step"]
print("$1", "step", round(d["ms_per_step"],3), {n:v for n,v in k.items() if n in ("k_acoustic_gather","k_dt_cellA","k_dt_cellB","k_dt_cellC<false>","k_dt_cellC<true>","k_dt_edge")}, d["check"]["combined_checksum"])
P
}
run v0 libmpas_b200_v0.so
run shipped libmpas_b200.so
run v1 libmpas_b200_v1.so
