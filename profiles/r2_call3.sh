#!/bin/bash
# round 2, GPU call 3: lane-pipeline acoustic kernel (tests + timing), column-alignment experiment, scalars parity
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "acoustic or scalars or level_counts or range_restricted" > gpurun_out/c3_pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest_a.log
tail -12 gpurun_out/c3_pytest_a.log
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 600 $B --acoustic 4 > gpurun_out/c3_ac4.json 2> gpurun_out/c3_ac4.err
timeout 600 $B --acoustic 2 > gpurun_out/c3_ac2.json 2> gpurun_out/c3_ac2.err
MPASB200_LP_ALIGN=16 timeout 600 $B --acoustic 4 > gpurun_out/c3_ac4_lp64.json 2> gpurun_out/c3_ac4_lp64.err
MPASB200_LP_ALIGN=16 MPASB200_CPB=8 timeout 600 $B --acoustic 4 > gpurun_out/c3_ac4_lp64_cpb8.json 2> gpurun_out/c3_ac4_lp64_cpb8.err
MPASB200_LP_ALIGN=16 MPASB200_CPB=8 timeout 600 $B --acoustic 2 > gpurun_out/c3_ac2_lp64_cpb8.json 2> gpurun_out/c3_ac2_lp64_cpb8.err
timeout 1200 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q > gpurun_out/c3_pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest_b.log
tail -8 gpurun_out/c3_pytest_b.log
P="python bench.py --mesh 163842 --steps 1 --warmup 3 --no-cpu --no-e2e --acoustic 4"
$P > gpurun_out/c3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_acoustic_lane" -s 4 -c 3 -o gpurun_out/prof_r2b $P > gpurun_out/c3_ncu.log 2>&1
