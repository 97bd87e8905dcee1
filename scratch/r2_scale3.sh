#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=gpurun_out/r2y
n=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 10 --no-cpu-baseline > ${P}_bench_n$n.json 2> ${P}_bench_n$n.err; echo "bench n=$n rc=$?"
python - <<PY
import json
d=json.loads(open('${P}_bench_n$n.json').read().strip().splitlines()[-1])
e=d['e2e']; m=d['multi']
print($n, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(e['value'],1), 'pcie/rank', round(e['pcie_gbs_per_rank'],1))
print('   multi ms', round(m['ms'],4), 'median', round(m['ms_median'],4), 'local', round(m['ms_local_only'],4), 'exposed', round(m['allreduce_ms_exposed'],4), 'bands', {k: round(v,3) for k,v in m['ms_by_bands'].items()}, 'ar alone', m.get('allreduce_alone_ms'))
PY
