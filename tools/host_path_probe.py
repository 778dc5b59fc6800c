"""What bounds the host-buffer step when several ranks share one host: the c2 step kernel launched back to
back (no host sync inside a block, CUDA events, max over ranks) with its actions and / or rewards in mapped host memory.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/host_path_probe.py

modes: dev (everything in HBM) | read (actions in pinned host memory) | write (rewards to pinned host memory) | both |
both_huge (the same with the host buffers in transparent huge pages, cudaHostRegister'ed) | ce_write (rewards to HBM, then
one copy-engine transfer of the 4 MB to the host per step)
"""
import ctypes
import json
import mmap
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def huge_pinned(nbytes: int):
    """(tensor of uint8, keepalive) over a 2 MB-aligned anonymous mapping advised MADV_HUGEPAGE and registered with CUDA."""
    two_mb = 2 << 20
    size = ((nbytes + two_mb - 1) // two_mb) * two_mb
    m = mmap.mmap(-1, size + two_mb, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(m))
    off = (-base) % two_mb
    try:
        m.madvise(mmap.MADV_HUGEPAGE, off, size)
    except Exception as e:  # noqa: BLE001
        print("madvise failed:", e, flush=True)
    t = torch.frombuffer(m, dtype=torch.uint8, count=size, offset=off)
    t.zero_()   # first touch
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), size, 1 | 2)   # portable | mapped
    assert int(rc) == 0, f"cudaHostRegister -> {rc}"
    return t, m


def anon_huge_kb() -> int:
    try:
        with open("/proc/self/smaps_rollup") as f:
            for line in f:
                if line.startswith("AnonHugePages"):
                    return int(line.split()[1])
    except Exception:  # noqa: BLE001
        pass
    return -1


def main():
    import bench
    from finenvs_b200 import parallel as par
    from finenvs_b200.data import loader

    rank, world, local_rank = par.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    N, W, steps, blocks = 1 << 20, 60, 20, 4
    prices, seg_start, seg_len, _ = bench.make_series("c2", W)
    series = loader.stage_series(prices, seg_start, seg_len, W, dev, torch.float32)
    env = par.make_sharded_env(N * world, rank, world, "probe", num_intervals=W, device_id=local_rank, series=series, seed=3,
                               random_reset="all", random_offset=True)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    a_dev = [torch.rand((N, 1), generator=g, device=dev) * 2 - 1 for _ in range(4)]
    a_pin = [a.cpu().pin_memory() for a in a_dev]
    obs = torch.empty((N, W, 5), dtype=torch.float32, device=dev)
    r_dev = torch.empty(N, dtype=torch.float32, device=dev)
    d_dev = torch.empty(N, dtype=torch.int32, device=dev)
    r_pin = torch.empty(N, dtype=torch.float32).pin_memory()
    thp_before = anon_huge_kb()
    keep = []
    try:
        a_huge = []
        for a in a_pin:
            t, m = huge_pinned(4 * N)
            keep.append(m)
            t = t[: 4 * N].view(torch.float32)
            t.copy_(a.view(-1))
            a_huge.append(t)
        t, m = huge_pinned(4 * N)
        keep.append(m)
        r_huge = t[: 4 * N].view(torch.float32)
        huge_ok = True
    except Exception as e:  # noqa: BLE001
        print("huge pages unavailable:", e, flush=True)
        huge_ok = False
    thp_after = anon_huge_kb()
    L, stream = env._L, env._stream()
    side = torch.cuda.Stream(device=dev)
    scratch = torch.zeros(N, dtype=torch.float32, device=dev)
    scratch_a = torch.zeros((N, 1), dtype=torch.float32, device=dev)

    def launch(a_ptr, r_ptr):
        env.step_count += 1
        rc = L.fe_step(env._pp, env._ps, env._pst, a_ptr, obs.data_ptr(), r_ptr, d_dev.data_ptr(), None, env.step_count, stream)
        assert rc == 0, rc

    def run(mode):
        def one(i):
            if mode == "dev":
                launch(a_dev[i % 4].data_ptr(), r_dev.data_ptr())
            elif mode == "read":
                launch(a_pin[i % 4].data_ptr(), r_dev.data_ptr())
            elif mode == "write":
                launch(a_dev[i % 4].data_ptr(), r_pin.data_ptr())
            elif mode == "both":
                launch(a_pin[i % 4].data_ptr(), r_pin.data_ptr())
            elif mode == "both_huge":
                launch(a_huge[i % 4].data_ptr(), r_huge.data_ptr())
            elif mode == "ce_write":
                launch(a_dev[i % 4].data_ptr(), r_dev.data_ptr())
                r_pin.copy_(r_dev, non_blocking=True)
            elif mode == "read+ce_async":   # zero-copy reads by the kernel, an INDEPENDENT 4 MB copy-engine write per step beside it
                launch(a_pin[i % 4].data_ptr(), r_dev.data_ptr())
                with torch.cuda.stream(side):
                    r_pin.copy_(scratch, non_blocking=True)
            elif mode == "dev+ce_async":    # no zero-copy traffic, the same independent copy-engine write per step
                launch(a_dev[i % 4].data_ptr(), r_dev.data_ptr())
                with torch.cuda.stream(side):
                    r_pin.copy_(scratch, non_blocking=True)
            elif mode == "dev+ce_h2d_async":   # an independent 4 MB copy-engine READ of host memory per step beside the kernel
                launch(a_dev[i % 4].data_ptr(), r_dev.data_ptr())
                with torch.cuda.stream(side):
                    scratch_a.copy_(a_pin[i % 4], non_blocking=True)
            elif mode == "write+ce_h2d_async":   # zero-copy writes by the kernel + independent copy-engine reads of host memory
                launch(a_dev[i % 4].data_ptr(), r_pin.data_ptr())
                with torch.cuda.stream(side):
                    scratch_a.copy_(a_pin[i % 4], non_blocking=True)
        for i in range(6):
            one(i)
        ms = []
        for _ in range(blocks):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                one(i)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / steps)
        v = torch.tensor(ms, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return sorted(v.tolist())[len(ms) // 2]

    out = {"world": world, "kernel": env.kernel_name(), "anon_huge_kb_before_after": [thp_before, thp_after]}
    for mode in ["dev", "read", "write", "both"] + (["both_huge"] if huge_ok else []) + ["ce_write", "read+ce_async", "dev+ce_async", "dev+ce_h2d_async", "write+ce_h2d_async", "both", "dev"]:
        ms = run(mode)
        out.setdefault(mode, []).append(round(ms, 4))

    # the real call (TimeSeriesEnv.step_host, host-synchronous, wall clock, max over ranks): torch-pinned buffers, then the same
    # env with every host buffer in huge pages
    import time

    def host_leg(actions, packed):
        def block():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                env.step_host(actions[i % 4], packed_dones=packed)
            return (time.perf_counter() - t0) * 1e3 / steps
        block()
        v = torch.tensor([block() for _ in range(blocks)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return round(sorted(v.tolist())[blocks // 2], 4)

    out["step_host_packed"] = [host_leg([a.view(-1) for a in a_pin], True)]
    out["step_host_int32"] = [host_leg([a.view(-1) for a in a_pin], False)]
    if huge_ok:
        bufs = list(env._host_bufs)
        for k, (nb, dt) in enumerate(((4 * N, torch.float32), (4 * N, torch.int32), (4 * ((N + 31) // 32), torch.uint8))):
            t, m = huge_pinned(nb)
            keep.append(m)
            bufs[3 + k] = t[:nb].view(dt)
        env._host_bufs = tuple(bufs)
        env._host_ptrs = tuple(t.data_ptr() for t in env._host_bufs)
        out["step_host_packed_huge"] = [host_leg(a_huge, True)]
        out["step_host_int32_huge"] = [host_leg(a_huge, False)]
        out["anon_huge_kb_end"] = anon_huge_kb()
    out["step_host_packed"].append(host_leg([a.view(-1) for a in a_pin], True) if not huge_ok else None)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
