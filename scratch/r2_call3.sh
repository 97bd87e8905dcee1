#!/bin/bash
set -o pipefail
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LI_B200_TRACE=1 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2g_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -5 gpurun_out/r2g_pytest_multi.log; grep "adj multi" gpurun_out/r2g_pytest_multi.log | sort | uniq -c | head -20
timeout 300 python scratch/multi_probe.py 64 > gpurun_out/probe.log 2>&1; grep -v "multi cta" gpurun_out/probe.log | tail -3
