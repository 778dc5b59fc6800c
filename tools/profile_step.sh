#!/bin/bash
# ncu launch list + one --set full capture of the step kernel for a bench workload (run under gpurun).
# usage: tools/profile_step.sh <workload> <tag> [kernel regex] [extra bench args]
WL=$1; TAG=$2; KRE=${3:-fe_gather}; EXTRA=$4
CMD="python bench.py --workload $WL --steps 6 --warmup 3 --blocks 1 --no-cpu-baseline --no-also $EXTRA"
timeout 120 $CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_l_${TAG}.log 2>&1
timeout 120 $CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 2 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_f_${TAG}.log
