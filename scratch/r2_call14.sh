#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2t_topo.txt 2>&1; head -14 gpurun_out/r2t_topo.txt
timeout 600 python -m pytest tests -q -m gpu -x -k "multi or nccl or comm" > gpurun_out/r2t_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -4 gpurun_out/r2t_pytest_multi.log
for flag in "" "--no-numa-bind"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline $flag > gpurun_out/r2t_bench2$flag.json 2> gpurun_out/r2t_bench2$flag.err; echo "bench2 $flag rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench2$flag.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['e2e']))
print(json.dumps(d.get('multi')))
PY
done
