#!/bin/bash
# round 2, GPU call 15 (8 GPUs): BASELINE config 3 (x1.163842 x 55) on 8 ranks with the final library -- the in-library distributed
# step with the sent-cell launches on the communication stream had only run on 2 ranks; checksum must equal N = 1 (68a7b531b35b42f9)
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c15_gpus.txt
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --mesh 163842"
timeout 600 $T8 > gpurun_out/c15_n8_163842.json 2> gpurun_out/c15_n8_163842.err
tail -c 400 gpurun_out/c15_n8_163842.json
python - <<P
import json
d=json.loads(open("gpurun_out/c15_n8_163842.json").read().strip().splitlines()[-1])
print("N=8 x1.163842: step", d["ms_per_step"], "value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), d["check"]["combined_checksum"])
P
