#!/bin/bash
cd /root/repo
for v in "-DFWD_WAVES_=3" "-DFWD_WAVES_=4" "-DFWD_WAVES_=2 -DFWD_NSTAGE_=2"; do
  touch linearinterpolators.jl_b200/csrc/li_sep2d.cu
  LI_NVCC_EXTRA="-DLI_DEV_FAST $v" python linearinterpolators.jl_b200/build.py > /tmp/b.log 2>&1 || { echo "build failed: $v"; tail -5 /tmp/b.log; continue; }
  echo "== variant: [$v]"
  timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value'],1), d['roofline']['fwd_ms'], d['roofline']['adj_ms'])"
done
