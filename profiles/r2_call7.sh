#!/bin/bash
# round 2, GPU call 7 (1 GPU): full GPU suite at HEAD, headline bench with both arms, launch list + ncu --set full of the gather kernels
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c7_pytest.log
tail -8 gpurun_out/c7_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c7_bench_headline.json 2> gpurun_out/c7_bench_headline.err
tail -c 600 gpurun_out/c7_bench_headline.json
P="python bench.py --mesh 163842 --steps 1 --warmup 3 --no-cpu --no-e2e"
timeout 300 $P > gpurun_out/c7_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/c7_launches.csv $P > gpurun_out/c7_ncu_list.log 2>&1
timeout 300 $P > gpurun_out/c7_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dt_edge$|k_dt_cellC|k_acoustic_gather|k_dt_theta_flux|k_divdamp|k_acoustic_lane|k_diag_cell|k_diag_edge" -s 36 -c 22 -o gpurun_out/prof_r2c $P > gpurun_out/c7_ncu_full.log 2>&1
ls -la gpurun_out
