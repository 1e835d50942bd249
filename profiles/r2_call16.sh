#!/bin/bash
# round 2, GPU call 16 (4 GPUs): BASELINE config 3 (x1.163842 x 55) on 4 ranks, final library; own-column L2 prefetch in k_dt_cellC (variant) on one GPU
set -x
cd "$GRAFT_REPO_ROOT"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-e2e --mesh 163842"
timeout 600 $T4 > gpurun_out/c16_n4_163842.json 2> gpurun_out/c16_n4_163842.err
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
run() {
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/$2 timeout 300 $B > gpurun_out/c16_$1.json 2> gpurun_out/c16_$1.err
}
run shipped libmpas_b200.so
run v3 libmpas_b200_v3.so
python - <<P
import json
d=json.loads(open("gpurun_out/c16_n4_163842.json").read().strip().splitlines()[-1])
print("N=4 x1.163842: step", d["ms_per_step"], "value", d["value"], d["check"]["combined_checksum"])
for t in ("shipped","v3"):
    d=json.loads(open(f"gpurun_out/c16_{t}.json").read().strip().splitlines()[-1]); k=d["kernels_ms_per_step"]
    print(t, "step", round(d["ms_per_step"],3), {n:v for n,v in k.items() if "cellC" in n}, d["check"]["combined_checksum"])
P
