#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scratch/fuzz_fused3d.py 300 1 2>&1 | tail -3
timeout 600 python scratch/fuzz_fused3d.py 200 2 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2ai_pytest.log
