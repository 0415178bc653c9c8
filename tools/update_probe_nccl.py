"""Layout-update latency over N ranks: rank 0 removes / places an item in its host mirror, every rank
gets the new layout. Compares the in-library path (jg_cde_broadcast: flatten + stream-ordered upload on the root, ONE
ncclBroadcast into a fixed-capacity exchange buffer, one host wait on the receivers) with round 1's path (two
torch.distributed broadcasts with a .item() in between + jg_layout_upload_device). Per update: time of the call on rank 0,
time of the call on the slowest receiver when all ranks enter together (barrier before), and the time until a small probe
batch enqueued right behind the update has its verdicts back (the layout is usable).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/update_probe_nccl.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jagua_rs_b200 as jg  # noqa: E402
from jagua_rs_b200 import multigpu  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(local_rank)
dist.init_process_group("nccl" if world > 1 else "gloo", device_id=torch.device("cuda", local_rank) if world > 1 else None,
                        **({} if world > 1 else dict(rank=0, world_size=1, init_method="tcp://127.0.0.1:29599")))
host_group = dist.new_group(backend="gloo")
name = sys.argv[1] if len(sys.argv) > 1 else "swim_half"
wl = jg.load_workload(name)
shapes = [jg.ItemShape(it.verts, it.poi, it.poles, it.piers) for it in wl.items]
if rank == 0:
    cde = jg.build_engine(wl, device=local_rank)
else:
    cde = jg.CDEngine(wl.cde_bbox, wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, device=local_rank)
    for i, sh in enumerate(shapes):
        cde.upload_item(i, sh)
bc = multigpu.LayoutBroadcaster(cde, dist, rank, world)
bc.broadcast(0)
probe_t = jg.PinnedBuffer((4096, 4), np.float32)
probe_t.array[:] = jg.make_queries(wl, 0, 4096, seed=3)
probe_o = jg.PinnedBuffer((4096,), np.uint8)
dyn = [(k, h) for k, h in zip(getattr(cde, "hazard_keys", []), wl.hazards) if h.dynamic] if rank == 0 else []
removed = []
rng = np.random.default_rng(1)


def mutate():
    """rank 0: remove a placed item or put a removed one back (Layout::remove_item / place_item on the host mirror)"""
    if removed and (len(removed) > 4 or rng.random() < 0.5):
        h = removed.pop()
        k = cde.register_hazard(h.kind, h.verts, h.poi, id=h.id, quality=h.quality, pk=h.pk, dynamic=True)
        dyn.append((k, h))
    else:
        k, h = dyn.pop(int(rng.integers(len(dyn))))
        cde.deregister_hazard_by_key(k)
        removed.append(h)


def run(mode, n=60):
    call, usable = [], []
    for it in range(n + 5):
        if rank == 0:
            mutate()
        cde.sync()
        dist.barrier(group=host_group)
        t0 = time.perf_counter()
        if mode == "library":
            bc.broadcast(0)
        else:
            blob = cde.flatten() if rank == 0 else None
            d_blob = multigpu.broadcast_blob(blob, dist, src=0, device=f"cuda:{local_rank}")
            cde.upload_layout_device(d_blob.data_ptr(), d_blob.numel(), epoch=it + 10)
            cde._dirty = False
        t1 = time.perf_counter()
        cde.detect_batch_async(0, probe_t.array, probe_o.array)
        cde.sync()
        t2 = time.perf_counter()
        if it >= 5:
            call.append(t1 - t0)
            usable.append(t2 - t0)
    tt = torch.tensor([np.median(call), np.percentile(call, 90), np.median(usable)], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(tt) for _ in range(world)]
    dist.all_gather(allr, tt) if world > 1 else allr.__setitem__(0, tt)
    a = torch.stack(allr).cpu().numpy() * 1e6
    return {"root_call_us_median": float(a[0, 0]), "root_call_us_p90": float(a[0, 1]),
            "slowest_receiver_call_us_median": float(a[1:, 0].max()) if world > 1 else None,
            "slowest_receiver_call_us_p90": float(a[1:, 1].max()) if world > 1 else None,
            "usable_us_median_slowest_rank": float(a[:, 2].max())}


res = {"workload": name, "n_gpus": world, "blob_bytes": int(len(cde.flatten())) if rank == 0 else None}
res["library_nccl"] = run("library")
if world > 1:
    res["round1_torch_two_broadcasts"] = run("torch")
if rank == 0:
    print(json.dumps(res))
cde.close()
dist.destroy_process_group()
