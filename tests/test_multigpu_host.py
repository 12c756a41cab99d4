"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous point sharding, the
post-kernel sum all_reduce and the value all_gather.  The per-rank evaluator is injected (the CPU
oracle) because there is no GPU here; on a GPU box the same code runs NCCL + the CUDA plan
(bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import itna_b200 as t
    import oracle as orc
    from itna_b200.parallel import evaluate_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = t.continuous_siteinds(t.named_grid((12, 1)), map_dimension=2)
    f = t.rand_itn(s, link_space=4, rng=3, normalise=True)
    packed = t.pack(f)
    pts = np.random.default_rng(5).random((1001, 2))  # not divisible by the world size
    ev = (lambda c, want_sum: ((None, complex(orc.evaluate(packed, c, orc.ORACLE_F64).sum(), 0.0)) if want_sum
                               else (orc.evaluate(packed, c, orc.ORACLE_F64), 0j)), False)
    full = evaluate_sharded(f, pts, [1, 2], evaluator=ev)
    mine = evaluate_sharded(f, pts, [1, 2], evaluator=ev, gather=False)
    total = evaluate_sharded(f, pts, [1, 2], evaluator=ev, reduce="sum")
    q.put((rank, full, mine, total))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from itna_b200.parallel import shard_bounds
    for n in (0, 1, 7, 1000, 10 ** 9 + 7):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= -(-n // world)


@pytest.mark.timeout(120)
def test_two_rank_gloo_shard_reduce_gather():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import itna_b200 as t
    import oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    s = t.continuous_siteinds(t.named_grid((12, 1)), map_dimension=2)
    f = t.rand_itn(s, link_space=4, rng=3, normalise=True)
    pts = np.random.default_rng(5).random((1001, 2))
    ref = orc.evaluate(t.pack(f), pts, orc.ORACLE_F64)
    for rank, full, mine, total in res:
        assert (full == ref).all()                      # gathered values: every rank has everything
        lo, hi = (0, 501) if rank == 0 else (501, 1001)
        assert (mine == ref[lo:hi]).all()               # contiguous blocks
        assert abs(total - ref.sum()) <= 1e-12 * np.abs(ref).sum()
    assert res[0][3] == res[1][3]                       # all_reduce: identical on all ranks
