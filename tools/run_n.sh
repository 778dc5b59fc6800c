#!/bin/bash
# Multi-GPU record runs (under gpurun --gpus N): bench c2, bench c4, ES rollout.  usage: tools/run_n.sh N TAG
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_c2_n$N.json 2> gpurun_out/${TAG}_bench_c2_n$N.err
$TR bench.py --gpus $N --steps 100 --warmup 5 --workload c4 > gpurun_out/${TAG}_bench_c4_n$N.json 2> gpurun_out/${TAG}_bench_c4_n$N.err
$TR tools/es_rollout.py --envs-per-gpu 524288 --steps 64 --generations 3 > gpurun_out/${TAG}_es_n$N.json 2> gpurun_out/${TAG}_es_n$N.err
tail -n 1 gpurun_out/${TAG}_bench_c2_n$N.json | cut -c1-400; tail -n 1 gpurun_out/${TAG}_bench_c4_n$N.json | cut -c1-300; tail -n 1 gpurun_out/${TAG}_es_n$N.json | cut -c1-600
tail -n 3 gpurun_out/${TAG}_es_n$N.err
