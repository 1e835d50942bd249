#!/bin/bash
# round 2, GPU call 12 (1 GPU): init_atm_case_jw on the device against the host generator
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "init_atm_case_jw" > gpurun_out/c12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c12_pytest.log
tail -30 gpurun_out/c12_pytest.log
