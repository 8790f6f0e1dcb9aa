"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): determinant space hash-partitioned over the
ranks, spawns exchanged with NCCL send/recv inside librimu_b200.so, results bit-exact against the single-rank
CPU oracle for integer walkers (replaces the reference's MPI tests, test/mpi_runtests.jl:140-190)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_worker(world, method, p2p, extra_env=None):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, RIMU_B200_METHOD=method, RIMU_B200_P2P=p2p)  # p2p=1: peer-direct exchange when CUDA IPC works
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("method,p2p", [("partition", "1"), ("partition", "0"), ("hash", "1")])
def test_two_rank_steps_match_oracle(built, method, p2p):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run_worker(2, method, p2p)
    assert out.count("mgpu ok") == (6 if (method, p2p) == ("partition", "1") else 4)  # + two initiator cases in direct mode


@pytest.mark.parametrize("world", [4, 8])
def test_many_rank_steps_match_oracle(built, world):
    """The same bit-exact comparison with the single-rank oracle at 4 and 8 ranks (direct exchange: W=1 and W=2 integer
    walkers, semistochastic Float64, two initiator cases) -- test/mpi_runtests.jl:41-155 runs its checks at the MPI size it is
    launched with; here the size is a parameter."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _run_worker(world, "partition", "1")
    assert out.count("mgpu ok") == 6
    assert f"world={world}" in out


@pytest.mark.parametrize("world", [2, 8])
def test_multi_rank_transcorrelated_energy(built, world):
    """Config 5's blocking-analysis energy check with the walker vector partitioned over `world` GPUs."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _run_worker(world, "partition", "1", {"RIMU_MGPU_MODE": "energy"})
    assert "mgpu energy" in out and '"within_5_sigma_plus_1pct": true' in out
    print(out[out.index("mgpu energy"):].splitlines()[0])
