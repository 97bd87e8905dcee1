#!/bin/bash
set -o pipefail
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2h_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -8 gpurun_out/r2h_pytest_all.log
LI_B200_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('multi'),indent=1))
PY
grep "adj multi" gpurun_out/r2h_bench.err | sort | uniq -c | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep2d_adj_multi -s 2 -c 1 -o gpurun_out/r2h_prof_multi \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 512 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
