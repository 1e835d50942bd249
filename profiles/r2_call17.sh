#!/bin/bash
# round 2, GPU call 17 (2 GPUs): launch timeline of one distributed step (both streams) on config 3's mesh and on the headline mesh
set -x
cd "$GRAFT_REPO_ROOT"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 400 $T --mesh 163842 --timeline gpurun_out/timeline_n2_163842.json > gpurun_out/c17_tl.log 2>&1; echo "rc=$?"
timeout 400 $T --mesh 163842 > gpurun_out/c17_n2_163842.json 2> gpurun_out/c17_n2_163842.err
python profiles/timeline_summary.py gpurun_out/timeline_n2_163842.json | head -30
python - <<P
import json
d=json.loads(open("gpurun_out/c17_n2_163842.json").read().strip().splitlines()[-1])
print("N=2 x1.163842: step", d["ms_per_step"], d["check"]["combined_checksum"])
P
