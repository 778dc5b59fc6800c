#!/bin/bash
# Time alternative builds of the library (finenvs_b200/libfe_*.so, built with -DFE_PIPE_* overrides) on c2.
for lib in finenvs_b200/libfe_*.so; do
  FINENVS_B200_LIB=$PWD/$lib timeout 120 python bench.py --workload ${WL:-c2} --variant pipe --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
done
