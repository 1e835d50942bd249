#!/bin/bash
# round 2, GPU call 29 (2 GPUs): the final library on two ranks -- the NCCL bitwise test (6 modes) and config 3's mesh at N = 2
set -x
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k nccl > gpurun_out/c29_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c29_pytest.log
tail -3 gpurun_out/c29_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 $T --mesh 163842 > gpurun_out/c29_n2_163842.json 2> gpurun_out/c29_n2_163842.err
python - <<P
import json
d=json.loads(open("gpurun_out/c29_n2_163842.json").read().strip().splitlines()[-1])
print("N=2 x1.163842: step", d["ms_per_step"], "value", d["value"], d["check"]["combined_checksum"])
P
