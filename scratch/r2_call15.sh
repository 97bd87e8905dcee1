#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench2.json 2> gpurun_out/r2u_bench2.err; echo "bench2 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2u_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d.get('multi')))
PY
tail -3 gpurun_out/r2u_bench2.err
