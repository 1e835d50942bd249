#!/bin/bash
# round 2, GPU call 27 (1 GPU): the final library -- full GPU suite, headline bench with both legs (x1.655362 x 55), launch list of one
# step and an ncu --set full capture of the dominant kernels at the headline size (traffic for bench.py's roofline)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c27_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c27_pytest.log
tail -4 gpurun_out/c27_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c27_bench_headline.json 2> gpurun_out/c27_bench_headline.err
tail -c 400 gpurun_out/c27_bench_headline.json
P="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
timeout 400 $P > gpurun_out/c27_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c27_launches.csv $P > gpurun_out/c27_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dt_edge$|k_dt_cellC|k_acoustic_lane|k_acoustic_gather|k_divdamp" -s 40 -c 12 -o gpurun_out/prof_r2f $P > gpurun_out/c27_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
