#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2x_pytest.log
timeout 600 python bench_configs.py --configs 4 2>&1 | tail -2 | cut -c1-220
timeout 600 python scratch/spectrum_probe.py 2>&1 | grep "1x\|1.5x\|0.5x\|Lanczos6"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --frames 512 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['fwd_ms'], d['roofline']['adj_ms'], d['multi']['ms'], d['multi']['us_per_frame_local'])"
