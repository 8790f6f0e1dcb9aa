"""bench.py host logic that must hold on every rank without a GPU: the steady-state detector and the reference arm's
rank handling (rank 0 alone runs and prints under torchrun; the others exit 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_settled_requires_flat_population_and_length():
    from bench import Settled
    # passing through the band while still growing is NOT settled: norm inside 5 % but drifting 4 % and the length exploding
    s = Settled(1000.0, 10)
    flags = [s.update(960.0 + 8.0 * i, 100 * (1.2 ** i)) for i in range(10)]
    assert not any(flags)
    # inside the band but the determinant count still changes by > 3 %
    s = Settled(1000.0, 10)
    assert not any(s.update(1000.0, 500 + 10 * i) for i in range(10))
    # flat walker number and length: settled exactly when the window is full
    s = Settled(1000.0, 10)
    flags = [s.update(1000.0 + (i % 3), 800 + (i % 2)) for i in range(12)]
    assert flags[:9] == [False] * 9 and all(flags[9:])
    # leaving the band resets nothing explicitly but the window must be clean again
    assert not s.update(1100.0, 800)
    assert not any(s.update(1000.0, 800) for _ in range(9))
    assert s.update(1000.0, 800)


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
                        "--cpu-walkers", "3000", "--equil", "10"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "spawn_attempts_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["n_gpus"] == 2 and line["steps"] == 2 and line["warmup"] == 1
