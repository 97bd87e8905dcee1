#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2v_pytest.log
