#!/bin/bash
# Record run at N GPUs (under gpurun --gpus N): the driver's two invocations, ours and the reference arm.  usage: tools/run_scale.sh N TAG
N=$1; TAG=$2
if [ "$N" = 1 ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "ours exit $?"
tail -n 1 gpurun_out/${TAG}_bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'] / 1e9, 3), 'G  ms', round(d['ms_per_step'], 4), ' e2e', round(d['e2e']['value'] / 1e9, 3), 'G  ms', round(d['e2e']['ms_per_step'], 4),
      ' int32', round(d['e2e']['int32_dones']['ms_per_step'], 4), ' also', {k: round(v['ms_per_step'], 4) for k, v in d.get('also', {}).items()})
print('collectives', d.get('collectives'))
print('clocks', d['clocks'], 'host', d.get('host'))
"
