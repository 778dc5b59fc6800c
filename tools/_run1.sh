set -x
python -m pytest tests/test_gpu_capture.py tests/test_gpu_es.py -x -q -m gpu 2>&1 | tail -15
ES="python tools/es_rollout.py --envs-per-gpu 524288 --steps 64 --generations 3"
$ES > gpurun_out/es3_fast_w8s3.json 2> gpurun_out/es3_fast.err; tail -c 600 gpurun_out/es3_fast.err
FE_ES_NO_FAST=1 $ES > gpurun_out/es3_generic.json 2>/dev/null
for v in w8s2 w12s2 w6s4; do FINENVS_B200_LIB=finenvs_b200/libfe_es_$v.so $ES > gpurun_out/es3_fast_$v.json 2>/dev/null; done
for f in gpurun_out/es3_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); g=d['generations'][-1]; print(g['ms_per_step'], g['env_steps_per_sec_with_policy'], g['mean_return'], g['theta_norm'])"; done
python tools/graph_rollout.py > gpurun_out/graph_c1.json 2> gpurun_out/graph_c1.err; tail -c 400 gpurun_out/graph_c1.err; cat gpurun_out/graph_c1.json
python tools/graph_rollout.py --envs 65536 > gpurun_out/graph_64k.json 2>/dev/null; cat gpurun_out/graph_64k.json
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1b_c2.json 2>gpurun_out/bench_r1b_c2.err; cat gpurun_out/bench_r1b_c2.json
