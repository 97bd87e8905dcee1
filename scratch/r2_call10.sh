#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2p_pytest.log
timeout 600 python bench_configs.py --configs 4 2>&1 | tail -2
timeout 600 python scratch/spectrum_probe.py 2>&1 | tail -30
