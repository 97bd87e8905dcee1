#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "every_size_pair" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2s_pytest.log
python scratch/bw_probe.py
timeout 600 python scratch/spectrum_probe.py 2>&1 | tail -12
