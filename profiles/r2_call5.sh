#!/bin/bash
# round 2, GPU call 5 (8 GPUs): strong scaling of the headline mesh with the in-library distributed step vs the host-side schedule
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c5_gpus.txt
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu"
timeout 900 $T8 > gpurun_out/c5_n8_native.json 2> gpurun_out/c5_n8_native.err
tail -c 400 gpurun_out/c5_n8_native.json
MPAS_B200_NATIVE_DIST=0 timeout 900 $T8 --no-e2e > gpurun_out/c5_n8_python.json 2> gpurun_out/c5_n8_python.err
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-e2e"
timeout 900 $T4 > gpurun_out/c5_n4_native.json 2> gpurun_out/c5_n4_native.err
