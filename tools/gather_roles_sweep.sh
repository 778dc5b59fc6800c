#!/bin/bash
for cfg in "24 2621440" "28 2246656" "32 1966080" "60 1048576"; do set -- $cfg
for lib in b6m12s3 b9m9s3 b12m6s3; do
FINENVS_B200_LIB=$PWD/finenvs_b200/libfe_ga_$lib.so timeout 120 python bench.py --workload c2 --window $1 --envs $2 --variant gather --steps 30 --warmup 5 --blocks 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().strip().splitlines() if x.startswith('{')]
if not l: print('W=$1 $lib: FAILED'); sys.exit()
d=json.loads(l[-1]); print('W=$1 N=$2 %-8s %.4f ms   e2e %.4f ms' % ('$lib', d['ms_per_step'], d['e2e']['ms_per_step']))"
done; done
