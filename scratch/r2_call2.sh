#!/bin/bash
# round 2, GPU call 2: the warp-specialised multi-frame kernel
set -o pipefail
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LI_B200_TRACE=1 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2f_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -30 gpurun_out/r2f_pytest_multi.log
LI_B200_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --frames 512 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('multi'),indent=1))
PY
grep "adj multi" gpurun_out/r2f_bench.err | sort | uniq -c | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep2d_adj_multi -s 2 -c 1 -o gpurun_out/r2f_prof_multi \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 64 > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2f_ncu.log
