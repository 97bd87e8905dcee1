#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for e in 0 1 0 1; do
  LI_B200_EXP=$e timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-multi > gpurun_out/r2k_bench_exp$e.json 2> gpurun_out/r2k_bench_exp$e.err; echo "bench exp=$e rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_bench_exp$e.json').read().strip().splitlines()[-1])
r=d['roofline']; n=d['notes']
print($e, d['ms_per_step'], 'fwd',r['fwd_ms'],'adj',r['adj_ms'],'eager',n.get('fwd_ms_eager_events'),n.get('adj_ms_eager_events'), 'frac', r['frac'])
PY
done
