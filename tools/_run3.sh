python -m pytest tests/test_gpu_es.py -x -q -m gpu 2>&1 | tail -5
ES="python tools/es_rollout.py --envs-per-gpu 524288 --steps 64 --generations 3"
$ES > gpurun_out/es5_w12s2.json 2> gpurun_out/es5.err; tail -c 300 gpurun_out/es5.err
for v in w16s1 w8s3; do FINENVS_B200_LIB=finenvs_b200/libfe_es_$v.so $ES > gpurun_out/es5_$v.json 2>/dev/null; done
for f in gpurun_out/es5_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); g=d['generations'][-1]; print(g['ms_per_step'], g['env_steps_per_sec_with_policy'], g['mean_return'], g['theta_norm'])"; done
ncu --set full --clock-control none --import-source on -k regex:fe_es_forward -s 70 -c 1 -f -o gpurun_out/prof_es_fast2 python tools/es_rollout.py --envs-per-gpu 524288 --steps 40 --generations 2 > gpurun_out/ncu_es_fast2.log 2>&1; tail -2 gpurun_out/ncu_es_fast2.log
