#!/bin/bash
# round 2, GPU call 10 (1 GPU): mesh-only init producers on the device (parity), tile kernel from the laboratory build (bit identity),
# k_dt_edge with in-thread L2 prefetch of its tail loads (variant library) against the shipped one
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "init_producers or edge_tiles or init_chain or smlstep" > gpurun_out/c10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest.log
tail -8 gpurun_out/c10_pytest.log
B="python bench.py --mesh 163842 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 $B > gpurun_out/c10_plain.json 2> gpurun_out/c10_plain.err
MPAS_B200_LIB=$GRAFT_REPO_ROOT/mpas_regent_b200/csrc/libmpas_b200_pf.so timeout 300 $B > gpurun_out/c10_pf.json 2> gpurun_out/c10_pf.err
python - <<P
import json
for t in ("plain","pf"):
    d=json.loads(open(f"gpurun_out/c10_{t}.json").read().strip().splitlines()[-1])
    k=d["kernels_ms_per_step"]
    print(t, "step", d["ms_per_step"], {n:v for n,v in k.items() if "dt_edge" in n}, d["check"]["combined_checksum"])
P
