#!/bin/bash
# round 2, GPU call 4 (2 GPUs): native distributed step (mpasb200_srk3_dist) -- bitwise check vs single partition, timing vs the host-side schedule
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c4_gpus.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "nccl_ranks" > gpurun_out/c4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c4_pytest.log
tail -30 gpurun_out/c4_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 600 $T --mesh 163842 > gpurun_out/c4_n2_163842_native.json 2> gpurun_out/c4_n2_163842_native.err
MPAS_B200_NATIVE_DIST=0 timeout 600 $T --mesh 163842 > gpurun_out/c4_n2_163842_python.json 2> gpurun_out/c4_n2_163842_python.err
timeout 600 python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/c4_n1_163842.json 2> gpurun_out/c4_n1_163842.err
timeout 900 $T > gpurun_out/c4_n2_655362_native.json 2> gpurun_out/c4_n2_655362_native.err
tail -c 300 gpurun_out/c4_n2_655362_native.json
