#!/bin/bash
# round 2, GPU call 8 (1 GPU): k_dt_edge reverted to the 56-register loop; tile-staged k_dt_edge_tile (TMA, de-duplicated columns) -- bit identity + timing
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "edge_tiles or task_parity" > gpurun_out/c8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pytest.log
tail -8 gpurun_out/c8_pytest.log
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for t in 0 16 216 8 408; do
  timeout 300 $B --edge-tiles $t > gpurun_out/c8_et$t.json 2> gpurun_out/c8_et$t.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c8_et$t.json").read().strip().splitlines()[-1])
k=d["kernels_ms_per_step"]
print("edge_tiles", $t, "step", d["ms_per_step"], {n:v for n,v in k.items() if "dt_edge" in n}, d["check"]["combined_checksum"])
P
done
