"""Host-side logic of the N > 1 path on CPU: gloo, world_size 2. The layout blob flattened on rank 0 reaches
rank 1 intact and is accepted by a host-side engine there; the batch shards cover every query exactly once."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import jagua_rs_b200 as jg
    from jagua_rs_b200 import multigpu

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = jg.load_workload("swim_half")
        blob = None
        if rank == 0:
            cde = jg.build_engine(wl, device=None)
            blob = cde.flatten()
        buf = multigpu.broadcast_blob(blob, dist, src=0)
        got = buf.numpy()
        # every rank can rebuild the same blob from the same hazards: the broadcast copy must equal it
        mine = jg.build_engine(wl, device=None).flatten()
        offs = np.concatenate([[0], np.cumsum(jg.split_batch(1000003, len(wl.items)))])
        lo, hi, ids, rel = multigpu.shard_segments(offs, rank, world)
        q.put((rank, hashlib.sha256(got.tobytes()).hexdigest(), hashlib.sha256(mine.tobytes()).hexdigest(), len(got), lo, hi,
               ids, rel.tolist()))
    finally:
        dist.destroy_process_group()


def test_layout_broadcast_and_sharding_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, h0, m0, n0, lo0, hi0, ids0, rel0), (r1, h1, m1, n1, lo1, hi1, ids1, rel1) = res
    assert h0 == h1 == m0 == m1 and n0 == n1 > 0      # the replicated layout is bit-identical everywhere
    assert lo0 == 0 and hi0 == lo1 and hi1 == 1000003  # contiguous, disjoint, complete
    assert rel0[-1] == hi0 - lo0 and rel1[-1] == hi1 - lo1
    assert ids0[-1] == ids1[0] or ids0[-1] + 1 == ids1[0]  # the cut falls inside or between segments


@pytest.mark.parametrize("n,world", [(0, 1), (1, 8), (7, 8), (1 << 20, 8), (100_000_003, 8), (10, 3)])
def test_shard_bounds_partition(n, world):
    sys.path.insert(0, ROOT)
    from jagua_rs_b200 import multigpu

    cuts = [multigpu.shard_bounds(n, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    sizes = [b - a for a, b in cuts]
    assert max(sizes) - min(sizes) <= 1
