#!/bin/bash
cd /root/repo
for v in "-DADS_OCC_=5" "-DADS_OCC_=6" "-DADS_OCC_=5 -DADS_WAVES_=2"; do
  touch linearinterpolators.jl_b200/csrc/li_sep2d_adjs.cuh
  LI_NVCC_EXTRA="-DLI_DEV_FAST $v" python linearinterpolators.jl_b200/build.py > /tmp/b.log 2>&1 || { echo "build failed: $v"; tail -5 /tmp/b.log; continue; }
  grep -i "spill" /tmp/b.log | head -2
  echo "== variant: [$v]"
  timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value'],1), d['roofline']['fwd_ms'], d['roofline']['adj_ms'])"
done
