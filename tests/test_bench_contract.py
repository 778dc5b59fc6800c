"""bench.py's JSON contract, checked on the arm that needs no GPU: `--impl reference` (the unmodified reference env from
baseline/_ref on the host cores, the C oracle port beside it).  The keys and their meaning are the driver's contract
(see bench.py's docstring)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--ref-widened-envs", "0", *extra], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["ms_per_step"] > 0 and d["value"] > 0
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["envs_per_gpu"] == 1 << 20   # the arm reports the GPU arm's config; the bounded sample is in cpu_baseline
    cb = d["cpu_baseline"]
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    assert cb["cpu_port"]["kind"] == "port" and cb["cpu_port"]["value"] > 0
    if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "finenvs")):
        assert cb["kind"] == "reference" and cb["tree"] == "baseline/_ref"
        assert {(r["envs"], r["threads"]) for r in cb["runs"]} >= {(1024, cb["cores"]), (1024, 1)}
    else:
        assert cb["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run("--gpus", "2", env=env) == []


def test_portfolio_workload_on_the_reference_arm():
    d = json.loads(_run("--workload", "c3")[0])
    assert d["config"]["assets"] == 30 and d["config"]["window"] == 128 and d["value"] > 0
    assert "workload c3" in d["cpu_baseline"]["cpu_port"]["sample"]


def test_reference_arm_forces_the_thread_count_under_torchrun():
    # torchrun exports OMP_NUM_THREADS=1; the arm must still use the cores the box has (round-1 SCALE ratios were void)
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    d = json.loads(_run("--gpus", "2", env=env)[0])
    ncpu = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncpu and d["cpu_baseline"]["cpu_port"]["cores"] == ncpu
    assert d["config"]["total_envs"] == 2 << 20
