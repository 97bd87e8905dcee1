#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_3d or sep3d_on or aligned or gather_adjoint_is_deterministic or (every_size_pair and 8-8) or (every_size_pair and 3-6)" > gpurun_out/r2ah_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2ah_memcheck.log; grep -m5 -A12 "Invalid\|misaligned" gpurun_out/r2ah_memcheck.log | head -40; tail -4 gpurun_out/r2ah_memcheck.log
