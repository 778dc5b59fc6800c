#!/bin/bash
# Sweep the tile kernel's envs-per-block / threads-per-block (tuning aid; run under gpurun).
for wl in c2 c4; do
for e in 4 8 16 32; do
for t in 32 64 128; do
  if [ $e -le $t ]; then
    FE_TILE_ENVS=$e FE_TILE_THREADS=$t python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl E=$e T=$t', round(d['ms_per_step'],4),'ms', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],4))"
  fi
done; done; done
