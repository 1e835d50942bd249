#!/bin/bash
# round 2, GPU call 1: full GPU test suite (incl. experimental chunked kernels), chunk_tiles experiment, headline bench with check
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt
free -g > gpurun_out/c1_host_mem.txt; nproc >> gpurun_out/c1_host_mem.txt
MPASB200_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
for ct in 0 4 16; do
  timeout 600 python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e --chunk-tiles $ct > gpurun_out/c1_bench163842_chunk$ct.json 2> gpurun_out/c1_bench163842_chunk$ct.err
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/c1_bench_headline.json 2> gpurun_out/c1_bench_headline.err
tail -c 600 gpurun_out/c1_bench_headline.json
