#!/bin/bash
# round 2, GPU call 14 (1 GPU): full GPU suite on the final kernels; packed statics in k_acoustic_gather (variant) against the shipped
# kernel; headline bench with both arms; ncu --set full of the dominant kernel at the headline size (traffic for bench.py's roofline)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c14_pytest.log
tail -4 gpurun_out/c14_pytest.log
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
run() {
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/$2 timeout 300 $B > gpurun_out/c14_$1.json 2> gpurun_out/c14_$1.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c14_$1.json").read().strip().splitlines()[-1])
k=d["kernels_ms_per_step"]
print("$1", "step", round(d["ms_per_step"],3), {n:v for n,v in k.items() if n in ("k_acoustic_gather","k_dt_edge")}, d["check"]["combined_checksum"])
P
}
run shipped libmpas_b200.so
run v2 libmpas_b200_v2.so
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c14_bench_headline.json 2> gpurun_out/c14_bench_headline.err
tail -c 300 gpurun_out/c14_bench_headline.json
P="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
timeout 400 $P > gpurun_out/c14_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dt_edge$|k_dt_cellC|k_acoustic_lane" -s 9 -c 6 -o gpurun_out/prof_r2d $P > gpurun_out/c14_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
