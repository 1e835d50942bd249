#!/bin/bash
# round 2, GPU call 21 (1 GPU): parity subset on the stall-guided kernels (incl. the new non-finite advection-coefficient test), then an
# ncu --set full capture with source of one launch of each heavy kernel on x1.163842 x 55 (the next stall profile)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "task_parity or full_step or nonfinite or vertical_mixing or kernel_forms or level_counts or range_restricted or emulated or ulp" > gpurun_out/c21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest.log
tail -5 gpurun_out/c21_pytest.log
P="python bench.py --mesh 163842 --steps 1 --warmup 3 --no-cpu --no-e2e"
timeout 300 $P > gpurun_out/c21_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dt_edge$|k_dt_cellC|k_acoustic_gather|k_dt_theta_flux|k_divdamp|k_smlstep|k_vert_imp|k_diag_cell|k_acoustic_lane" -s 40 -c 26 -o gpurun_out/prof_r2e $P > gpurun_out/c21_ncu.log 2>&1
tail -3 gpurun_out/c21_ncu.log; ls -la gpurun_out | tail -4
