python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "step_host" 2>&1 | tail -4
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline"
pick() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', 'dev ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'e2e G', round(d['e2e']['value']/1e9,3))"; }
$B > gpurun_out/b6_c2_zc.json 2>gpurun_out/b6.err; tail -c 300 gpurun_out/b6.err; pick gpurun_out/b6_c2_zc.json
$B --workload c4 > gpurun_out/b6_c4_zc.json 2>/dev/null; pick gpurun_out/b6_c4_zc.json
$B --workload c3 > gpurun_out/b6_c3_zc.json 2>/dev/null; pick gpurun_out/b6_c3_zc.json
for c in 4 8 12; do FE_HOST_NO_ZEROCOPY=1 FE_HOST_CHUNKS=$c $B > gpurun_out/b6_c2_ch$c.json 2>/dev/null; pick gpurun_out/b6_c2_ch$c.json; done
python - <<'PY'
import torch
x=torch.empty(1258291200//4, dtype=torch.float32, device='cuda')
y=torch.empty_like(x)
def t(f,n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
ms=t(lambda: x.zero_()); print("fill 1.258 GB: %.4f ms  %.0f GB/s"%(ms, 1.2583/ms*1e3))
ms=t(lambda: x.fill_(1.5)); print("fill_ 1.258 GB: %.4f ms  %.0f GB/s"%(ms, 1.2583/ms*1e3))
ms=t(lambda: y.copy_(x)); print("copy 1.258 GB: %.4f ms  %.0f GB/s (r+w)"%(ms, 2*1.2583/ms*1e3))
s=x.sum(); ms=t(lambda: x.sum()); print("read-sum 1.258 GB: %.4f ms  %.0f GB/s"%(ms, 1.2583/ms*1e3))
PY
