#!/bin/bash
# Scaling evidence on one 8-GPU node: bench.py (config #2 weak + config #5 strong block) at N = 8, 4, 2; config #3 sharded.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=gpurun_out/r2y
nvidia-smi topo -m > ${P}_topo.txt 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 10 --no-cpu-baseline > ${P}_bench_n$n.json 2> ${P}_bench_n$n.err; echo "bench n=$n rc=$?"
  python - <<PY
import json
d=json.loads(open('${P}_bench_n$n.json').read().strip().splitlines()[-1])
e=d['e2e']; m=d['multi']
print($n, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(e['value'],1), 'pcie/rank', round(e['pcie_gbs_per_rank'],1), e['host_numa_binding'])
print('   multi ms', m.get('ms'), 'local', m.get('ms_local_only'), 'exposed', m.get('allreduce_ms_exposed'), 'bands', m.get('ms_by_bands'), 'ar alone', m.get('allreduce_alone_ms'))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu-baseline --no-multi --no-numa-bind > ${P}_bench_n8_nobind.json 2> /dev/null; echo "bench n=8 nobind rc=$?"
python -c "
import json
d=json.loads(open('${P}_bench_n8_nobind.json').read().strip().splitlines()[-1]); e=d['e2e']
print('nobind e2e', round(e['value'],1), 'pcie/rank', round(e['pcie_gbs_per_rank'],1))"
for n in 8 1; do
  if [ $n = 1 ]; then timeout 900 python bench_multi.py --configs 3 > ${P}_multi3_n$n.jsonl 2> ${P}_multi3_n$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench_multi.py --configs 3 > ${P}_multi3_n$n.jsonl 2> ${P}_multi3_n$n.err; fi
  echo "multi3 n=$n rc=$?"; cut -c1-300 ${P}_multi3_n$n.jsonl
done
