for a in 0 1 2 3 4; do echo -n "ablate=$a "; LI_B200_ABLATE=$a python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"adj_ms": [0-9.]*'; done
