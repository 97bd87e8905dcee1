#!/bin/bash
# round 2, GPU call 1: new tests first (fail fast), then the whole suite, the bench, and one ncu capture
set -o pipefail
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2a_pytest_multi.log 2>&1; echo "pytest multi rc=$?" 
tail -30 gpurun_out/r2a_pytest_multi.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -15 gpurun_out/r2a_pytest_all.log
LI_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json; grep "adj multi" gpurun_out/r2a_bench.err | sort | uniq -c | head
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 64 > gpurun_out/r2a_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sep2d_adj_multi -s 2 -c 2 -o gpurun_out/r2a_prof_multi \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 64 > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2a_ncu.log
