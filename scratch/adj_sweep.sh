#!/bin/bash
cd /root/repo
touch linearinterpolators.jl_b200/csrc/li_sep2d_adjs.cuh
LI_NVCC_EXTRA="-DLI_DEV_FAST -DADS_ABLATE" python linearinterpolators.jl_b200/build.py > /tmp/b.log 2>&1 || { echo "build failed"; tail -5 /tmp/b.log; }
timeout 120 python scratch/gap_probe.py 2>&1 | grep "adj:"
LI_B200_ADJ_DEBUG=4 timeout 120 python scratch/gap_probe.py 2>&1 | grep "adj:"
