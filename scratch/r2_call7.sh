#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LI_B200_TRACE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_3d and up2 and float32" 2>&1 | grep "li_b200\|passed\|failed" | tail -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sep2d_fwd_kernel|sep2d_adj_stream" -s 2 -c 2 -o gpurun_out/r2m_3d python bench_configs.py --configs 4 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2m_ncu.log
