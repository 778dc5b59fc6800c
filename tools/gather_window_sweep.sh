#!/bin/bash
# gather vs pipe vs tile across the windows `auto` gives to the gather kernel (legal TMA rows, 24 <= W <= 100 for f32), the same
# ~1.26 GB of observations per step; one line per (window, variant).  Run under gpurun.
for cfg in "24 2621440" "28 2246656" "32 1966080" "40 1572864" "48 1310720" "60 1048576" "80 786432" "100 629120"; do set -- $cfg
for v in gather pipe tile; do
timeout 120 python bench.py --workload c2 --window $1 --envs $2 --variant $v --steps 30 --warmup 5 --blocks 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().strip().splitlines() if x.startswith('{')]
if not l: print('W=$1 N=$2 $v: FAILED'); sys.exit()
d=json.loads(l[-1]); r=d['roofline']; print('W=$1 N=$2 %-6s %-32s %.4f ms   e2e %.4f ms' % ('$v', r['kernel'], d['ms_per_step'], d['e2e']['ms_per_step']))"
done; done
