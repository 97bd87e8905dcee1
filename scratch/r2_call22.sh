#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for rep in 1 2; do
for v in "" _f6 _f8; do
  LI_B200_LIB=$PWD/linearinterpolators.jl_b200/libli_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-multi 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('variant \"$v\"', 'step', round(d['ms_per_step']*1e3,2), 'fwd', round(r['fwd_ms']*1e3,2), 'adj', round(r['adj_ms']*1e3,2))"
done; done
for v in "" _f6 _f8; do
  echo "spectrum $v"; LI_B200_LIB=$PWD/linearinterpolators.jl_b200/libli_b200$v.so timeout 300 python scratch/spectrum_probe.py 2>/dev/null | awk '{print $1,$2,$3,$5,$6}' | head -9
  LI_B200_LIB=$PWD/linearinterpolators.jl_b200/libli_b200$v.so timeout 300 python bench_configs.py --configs 4 2>/dev/null | head -1 | cut -c100-200
done
