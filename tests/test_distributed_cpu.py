"""CPU tests of the multi-process host logic (gloo, world_size 2): rendezvous on 127.0.0.1, id
broadcast, slab ownership, gather of per-rank slabs into the reference's LU4 order."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from lbmpkg import lbm
    import cfd_codes_b200.distributed as D
    dist = D.init_process_group("gloo")
    nid = D.broadcast_nccl_id(make_id=lambda: bytes(range(128)))
    k0, k1 = lbm.slab_range(N, rank, world)
    # synthetic "distributions": value encodes (q, k, j, i) so any misplacement shows
    full = np.arange(19 * N ** 3, dtype=np.float64).reshape(19, N, N * N)
    mine = full[:, k0:k1].copy()
    got = D.gather_slabs(mine, N, N * N, 19)
    mx = D.allreduce_max(rank + 0.5)
    ok = (nid == bytes(range(128))) and mx == world - 0.5
    if rank == 0:
        ok = ok and np.array_equal(got, full)
    q.put((rank, bool(ok), (k0, k1)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 12), (3, 13)])
def test_gloo_slab_gather(world, N):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    spans = sorted(s for _, _, s in res)
    assert spans[0][0] == 0 and spans[-1][1] == N and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
