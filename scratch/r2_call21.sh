#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for n in 8 4; do
for e in 0 8; do
LI_B200_EXP=$e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$e bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ab_bench${n}_e$e.json 2> gpurun_out/r2ab_bench${n}_e$e.err; echo "bench n=$n exp=$e rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ab_bench${n}_e$e.json').read().strip().splitlines()[-1]); m=d['multi']
print('n',$n,'exp',$e,'ms',round(m['ms'],4),'median',round(m['ms_median'],4),'local',round(m['ms_local_only'],4),'exposed',round(m['allreduce_ms_exposed'],4),'bands',m['ms_by_bands'],'alone',m['allreduce_alone_ms'], 'finite', m['results_finite'])
PY
done; done
