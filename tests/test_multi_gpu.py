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


@pytest.mark.parametrize("method,p2p", [("partition", "1"), ("partition", "0"), ("hash", "1")])
def test_two_rank_steps_match_oracle(built, method, p2p):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, RIMU_B200_METHOD=method, RIMU_B200_P2P=p2p)  # p2p=1: peer-direct exchange when CUDA IPC works
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("mgpu ok") == (6 if (method, p2p) == ("partition", "1") else 4)  # + two initiator cases in direct mode
