#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "3d or 3_d" > gpurun_out/r2n_pytest3d.log 2>&1; echo "pytest 3d rc=$?"
tail -15 gpurun_out/r2n_pytest3d.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "config4" > gpurun_out/r2n_pytest_cfg4.log 2>&1; echo "pytest cfg4 rc=$?"
tail -3 gpurun_out/r2n_pytest_cfg4.log
timeout 600 python bench_configs.py --configs 4 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sep2d_adj_stream" -s 1 -c 1 -o gpurun_out/r2n_3d_adj python bench_configs.py --configs 4 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sep2d_fwd" -s 1 -c 1 -o gpurun_out/r2n_3d_fwd python bench_configs.py --configs 4 > gpurun_out/r2n_ncu2.log 2>&1; echo "ncu rc=$?"
