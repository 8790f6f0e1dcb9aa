"""Multi-rank host logic on CPU (gloo, world_size 2): determinant space is partitioned by address hash,
each rank steps only the parents it owns, spawned records are routed to the owner of the child
address (all_to_all) and annihilated there.  Because every random draw is keyed on (step key, address,
attempt), the union of the ranks' results must equal the single-rank step BIT-EXACTLY for integer
walkers -- the property the NCCL path relies on (communicators.jl:77-81,546-606 in the reference).
The compute in this test is the oracle (the GPU exchange itself is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    from oracle import oracle as orc
    import rimu_b200 as R
    from rimu_b200 import _lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _lib.lib()
    onr = (1,) * 8
    oh = orc.OracleHam("HubbardReal1D", "bose", onr, u=6.0, t=1.0)

    def owner(key):
        kk = np.ascontiguousarray(np.array(key, dtype=np.uint64))
        return L.rimu_addr_owner(kk.ctypes.data_as(_lib._u64p), len(kk), world)  # the PRODUCT's partition function

    # global reference run on every rank (cheap) and partitioned run
    gk, gv = np.array([oh.start_key], dtype=np.uint64), np.array([3000], dtype=np.int64)
    lk = gk[[owner(k) == rank for k in gk]] if len(gk) else gk
    lv = gv[[owner(k) == rank for k in gk]]
    ok = True
    for step in range(5):
        p = orc.make_params(orc.STYLE_INTEGER, shift=1.0, dtau=0.01, key=orc.step_key(11, step))
        gk, gv, gst = oh.step(p, gk, gv)
        # local parents -> per-parent columns (un-annihilated records are what crosses the wire)
        send = [[] for _ in range(world)]
        for k, v in zip(lk, lv):
            ck, cv, _ = oh.step(p, k.reshape(1, -1), np.array([v]))
            for kk, vv in zip(ck, cv):
                send[owner(kk)].append((tuple(int(x) for x in kk), int(vv)))
        recv = [None] * world
        dist.all_to_all_object_list(recv, send) if hasattr(dist, "all_to_all_object_list") else None
        if recv[0] is None:  # older torch: emulate with all_gather_object
            allsend = [None] * world
            dist.all_gather_object(allsend, send)
            recv = [allsend[src][rank] for src in range(world)]
        recs = [r for part in recv for r in part]
        if recs:
            kk = np.array([r[0] for r in recs], dtype=np.uint64).reshape(len(recs), -1)
            vv = np.array([r[1] for r in recs], dtype=np.int64)
            lk, lv = orc.annihilate(oh.W, kk, vv)
        else:
            lk, lv = np.zeros((0, oh.W), dtype=np.uint64), np.zeros(0, dtype=np.int64)
        # all ranks own disjoint key sets whose union is the global vector
        mine = np.array([owner(k) == rank for k in gk], dtype=bool)
        ok &= np.array_equal(lk, gk[mine]) and np.array_equal(lv, gv[mine])
        # walkernumber/length all-reduce (pdvec.jl:896-902)
        t = torch.tensor([float(np.abs(lv).sum()), float(len(lv))], dtype=torch.float64)
        dist.all_reduce(t)
        ok &= t[0].item() == float(np.abs(gv).sum()) and int(t[1].item()) == len(gv)
    q.put((rank, bool(ok), int(len(lk))))
    dist.barrier()
    dist.destroy_process_group()


def test_hash_partitioned_step_equals_global_step(built):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True], res
    assert all(r[2] > 0 for r in res), "both ranks must own part of the vector"
