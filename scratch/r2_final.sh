#!/bin/bash
# Evidence of the round: bench line, launch list, ncu captures of the headline / 3-D / multi-frame kernels (exported to csv on
# the box: the reports themselves exceed what gpurun copies back), the other configs.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=gpurun_out/r2z
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${P}_smoke.log
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
timeout 600 python bench_configs.py --batch3 8 --frames5 128 > ${P}_configs.log 2>&1; echo "configs rc=$?"
timeout 600 python scratch/spectrum_probe.py > ${P}_spectrum.log 2>&1; echo "spectrum rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-multi > ${P}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/${name} "$@" > ${P}_ncu_${name}.log 2>&1; echo "ncu $name rc=$?"
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > ${P}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > ${P}_${name}_src.csv.gz
  rm -f /tmp/${name}.ncu-rep
}
cap headline "sep2d_fwd_kernel|sep2d_adj_stream" 8 2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-multi
cap 3d_fwd "sep2d_fwd" 1 1 python bench_configs.py --configs 4
cap 3d_adj "sep2d_adj_stream" 1 1 python bench_configs.py --configs 4
cap multi "sep2d_adj_multi" 2 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 512
ls -la gpurun_out/r2z*; du -sh gpurun_out
