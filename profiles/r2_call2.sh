#!/bin/bash
# round 2, GPU call 2: GPU test suite (no -x), gather_stage / acoustic ablations on x1.163842, ncu of the new kernels
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
for gs in 0 1 2 4 -1; do
  timeout 600 $B --gather-stage $gs --acoustic 3 > gpurun_out/c2_gs${gs}_ac3.json 2> gpurun_out/c2_gs${gs}_ac3.err
done
timeout 600 $B --gather-stage 0 --acoustic 2 > gpurun_out/c2_gs0_ac2.json 2> gpurun_out/c2_gs0_ac2.err
timeout 600 $B --gather-stage 0 --acoustic 3 --acoustic-cols 8 > gpurun_out/c2_gs0_ac3_c8.json 2> gpurun_out/c2_gs0_ac3_c8.err
timeout 600 $B --gather-stage 0 --acoustic 3 --acoustic-cols 2 > gpurun_out/c2_gs0_ac3_c2.json 2> gpurun_out/c2_gs0_ac3_c2.err
P="python bench.py --mesh 163842 --steps 1 --warmup 3 --no-cpu --no-e2e"
$P > gpurun_out/c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_dt_edge_s|k_acoustic_seq|k_acoustic_gather_s|k_dt_theta_flux_s|k_dt_cellC" -s 60 -c 12 -o gpurun_out/prof_r2a $P > gpurun_out/c2_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
