#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LI_B200_TRACE=1 timeout 300 python -m pytest tests/test_multi_gpu.py -q -m gpu -x > gpurun_out/r2aa_pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"
grep "li_b200\]" gpurun_out/r2aa_pytest_2gpu.log | sort | uniq -c | head; tail -5 gpurun_out/r2aa_pytest_2gpu.log
for e in 0 8; do
LI_B200_EXP=$e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$e bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2aa_bench2_e$e.json 2> gpurun_out/r2aa_bench2_e$e.err; echo "bench2 exp=$e rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_bench2_e$e.json').read().strip().splitlines()[-1]); m=d['multi']
print('exp',$e,'ms',m['ms'],'local',m['ms_local_only'],'exposed',m['allreduce_ms_exposed'],'bands',m['ms_by_bands'],'alone',m['allreduce_alone_ms'])
PY
done
