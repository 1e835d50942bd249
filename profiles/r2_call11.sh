#!/bin/bash
# round 2, GPU call 11 (1 GPU): init producers + tile kernel (lab) + kernel_forms tests; k_dt_cellC as two launches at three register caps
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "init_producers or edge_tiles or init_chain or kernel_forms or staged" > gpurun_out/c11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c11_pytest.log
tail -8 gpurun_out/c11_pytest.log
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
run() { # tag, lib, extra args
  MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/$2 timeout 300 $B $3 > gpurun_out/c11_$1.json 2> gpurun_out/c11_$1.err
  python - <<P
import json
d=json.loads(open("gpurun_out/c11_$1.json").read().strip().splitlines()[-1])
k=d["kernels_ms_per_step"]
print("$1", "step", d["ms_per_step"], {n:v for n,v in k.items() if "cellC" in n}, d["check"]["combined_checksum"])
P
}
run fused libmpas_b200.so "--kernel-forms 0"
run split5 libmpas_b200.so "--kernel-forms 1"
run split4 libmpas_b200_s4.so "--kernel-forms 1"
run split6 libmpas_b200_s6.so "--kernel-forms 1"
