#!/bin/bash
# auto-variant check across windows: ms per step and algorithmic GB/s for pipe / tile / auto (run under gpurun)
for cfg in "4 4194304" "16 2097152" "60 1048576" "128 524288" "390 131072" "1024 65536"; do set -- $cfg
for v in auto pipe tile; do
timeout 120 python bench.py --workload ${WL:-c2} --window $1 --envs $2 --variant $v --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=sys.stdin.read().strip().splitlines()
if not l: print('W=$1 N=$2 $v: FAILED'); sys.exit()
d=json.loads(l[-1]); r=d['roofline']; print('W=$1 N=$2 %-5s %-32s %.4f ms  %.0f GB/s alg (%.2f)  e2e %.4f ms' % ('$v', r['kernel'], d['ms_per_step'], r['achieved'], r['frac'], d['e2e']['ms_per_step']))"
done; done
