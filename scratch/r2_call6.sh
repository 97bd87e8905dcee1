#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "3d or 3_d" > gpurun_out/r2l_pytest3d.log 2>&1; echo "pytest 3d rc=$?"
tail -30 gpurun_out/r2l_pytest3d.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "config4" > gpurun_out/r2l_pytest_cfg4.log 2>&1; echo "pytest cfg4 rc=$?"
tail -5 gpurun_out/r2l_pytest_cfg4.log
LI_B200_TRACE=1 timeout 600 python bench_configs.py --configs 4 2>&1 | tail -12
LI_B200_EXP=2 timeout 600 python bench_configs.py --configs 4 2>&1 | tail -4
