#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2ad_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/r2ad_bench.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r2ad_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value', d['value'], 'step', d['ms_per_step'], 'fwd', r['fwd_ms'], 'adj', r['adj_ms'], 'frac', r['frac'], r['traffic'], r['traffic_captured_at'], 'e2e', d['e2e']['value'], 'multi', d['multi']['ms'], 'cpu', d['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
