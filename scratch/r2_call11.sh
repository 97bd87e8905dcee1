#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -m gpu -x -k "3d or 3_d or config4" > gpurun_out/r2q_pytest3d.log 2>&1; echo "pytest 3d rc=$?"
tail -4 gpurun_out/r2q_pytest3d.log
for v in "" _b _c _d; do
  echo "== variant '$v'"
  LI_B200_LIB=$PWD/linearinterpolators.jl_b200/libli_b200$v.so timeout 600 python bench_configs.py --configs 4 2>&1 | tail -2 | cut -c1-200
done
for v in "" _c; do
  echo "== spectrum variant '$v'"
  LI_B200_LIB=$PWD/linearinterpolators.jl_b200/libli_b200$v.so timeout 600 python scratch/spectrum_probe.py 2>&1 | grep "1x\|1.5x\|0.5x"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sep2d_fwd" -s 1 -c 1 -o gpurun_out/r2q_3d_fwd python bench_configs.py --configs 4 > gpurun_out/r2q_ncu2.log 2>&1; echo "ncu rc=$?"
