#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for rep in 1 2; do
for v in "" _w3 _w4; do
  LI_B200_LIB=$PWD/linearinterpolators.jl_b200/libli_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-multi 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('variant \"$v\"', 'step', round(d['ms_per_step']*1e3,2), 'fwd', round(r['fwd_ms']*1e3,2), 'adj', round(r['adj_ms']*1e3,2))"
done; done
