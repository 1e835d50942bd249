#!/bin/bash
# round 2, GPU call 6 (2 GPUs): sent-cell launches on the communication stream -- bitwise check (6 modes) + timing on config 3's mesh
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "nccl_ranks" > gpurun_out/c6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest.log
tail -20 gpurun_out/c6_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 600 $T --mesh 163842 > gpurun_out/c6_n2_163842_native.json 2> gpurun_out/c6_n2_163842_native.err
timeout 600 $T --mesh 40962 > gpurun_out/c6_n2_40962_native.json 2> gpurun_out/c6_n2_40962_native.err
MPAS_B200_NATIVE_DIST=0 timeout 600 $T --mesh 40962 > gpurun_out/c6_n2_40962_python.json 2> gpurun_out/c6_n2_40962_python.err
