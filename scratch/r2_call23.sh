#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -q -m gpu -x 2>&1 | grep -v "li_b200\]" | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_multi_gpu.py > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2ac_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2ac_bench.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value', d['value'], 'step', d['ms_per_step'], 'fwd', r['fwd_ms'], 'adj', r['adj_ms'], 'frac', r['frac'], 'e2e', d['e2e']['value'], 'multi', d['multi']['ms'])"
