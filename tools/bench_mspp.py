"""BASELINE.json config 5: multi-strip packing with a minimum separation distance - a long detect sweep over the
strips of one instance, sharded over the GPUs of the box, with a layout update between sub-batches.

    python tools/bench_mspp.py [--instance jakobs1] [--queries-per-gpu 12500000] [--sub-batch 1048576]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_mspp.py --gpus N ...

Layouts: workloads/<instance>_ms<k>.npz, one per strip (tools/make_workloads.py: synthesised ExtMSPStrip, items
inflated / strips deflated by half the separation with the repo's own offsetter - the reference ships no MSPP
asset, optimizer or bench). Every rank owns one engine per strip (layouts are small and replicated) and its own
seeded slice of the sweep (weak scaling; no collective on the query path). Before every sub-batch rank 0
applies one layout update to one strip - Layout::place_item of an item at a transform the previous sub-batch
reported free, or Layout::remove_item of an earlier placement (entities/layout.rs:40-74) - flattens that
strip's quadtree and broadcasts the blob over NCCL INSIDE the library (jg_cde_broadcast: one ncclBroadcast into a
fixed-capacity exchange buffer and one host wait per update on the receiving ranks, none on rank 0). The reported
queries/s include those stalls. After the timed sweep rank 0 replays the updates on the CPU oracle and compares
the verdicts of every sub-batch it resolved (mismatches must be 0).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT,):
    sys.path.insert(0, p)
import jagua_rs_b200 as jg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--instance", default="jakobs1")
ap.add_argument("--queries-per-gpu", type=int, default=12_500_000)
ap.add_argument("--sub-batch", type=int, default=1 << 20)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--lookahead", type=int, default=0, choices=[0, 1],
                help="1: the update before sub-batch b is decided from the verdicts of sub-batch b-2 (waited for through a "
                     "fence), so rank 0 never waits for the sub-batch in flight; 0: from sub-batch b-1 (fully dependent)")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
from jagua_rs_b200 import multigpu  # noqa: E402
numa = multigpu.bind_to_gpu_numa_node(local_rank) if world > 1 else None
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

strips = []
k = 0
while os.path.exists(os.path.join(ROOT, "workloads", f"{args.instance}_ms{k}.npz")):
    strips.append(jg.load_workload(f"{args.instance}_ms{k}"))
    k += 1
assert strips, f"no workloads/{args.instance}_ms*.npz"
S = len(strips)
n_items = len(strips[0].items)
shapes = [jg.ItemShape(it.verts, it.poi, it.poles, it.piers) for it in strips[0].items]

# one engine per strip; rank 0 builds + flattens, the others receive the blob (one NCCL communicator per strip engine)
cdes, bcs = [], []
for s, wl in enumerate(strips):
    if rank == 0:
        cde = jg.build_engine(wl, device=local_rank)
    else:
        cde = jg.CDEngine(wl.cde_bbox, wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, device=local_rank)
        for i, sh in enumerate(shapes):
            cde.upload_item(i, sh)
    if world > 1:
        bcs.append(multigpu.LayoutBroadcaster(cde, dist, rank, world))
        bcs[-1].broadcast(root=0)
    else:
        cde.sync_layout()
    cdes.append(cde)

# the sweep: sub-batches of `sub_batch` queries, split evenly over strips and (inside a strip) over item types
n_sub = max(1, args.queries_per_gpu // args.sub_batch)
per_strip = args.sub_batch // S
sizes = jg.split_batch(per_strip, n_items)
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
ids = list(range(n_items))
# two seeded query sets per strip, alternated (generating 100 M distinct host records is not what is measured)
N_SETS = 2
host_t = [[jg.PinnedBuffer((per_strip, 4), np.float32) for _ in range(N_SETS)] for _ in range(S)]
for s in range(S):
    for v in range(N_SETS):
        for i, n in enumerate(sizes):
            host_t[s][v].array[offs[i]:offs[i + 1]] = jg.make_queries(strips[s], i, n, seed=1000 * rank + 10 * v + s)
keep = rank == 0 and not args.no_check
host_out = [[jg.PinnedBuffer((per_strip,), np.uint8) for _ in range(n_sub if keep else 2)] for _ in range(S)]


def barrier():
    for c in cdes:
        c.sync()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


fences = {}  # (sub-batch, strip) -> fence token of that strip's engine (rank 0, lookahead mode)


def run_sub(b, out_slot):
    for s in range(S):
        cdes[s].detect_multi(ids, offs, host_t[s][b % N_SETS].array, out=host_out[s][out_slot].array, wait=False)
        if args.lookahead and rank == 0:
            fences[(b, s)] = cdes[s].fence()


rng = np.random.default_rng(7)
updates = []   # (sub-batch index, strip, op, item, transform4 / None, pk) applied BEFORE that sub-batch
placed = [[] for _ in range(S)]  # PItemKeys placed by the sweep, per strip
next_pk = 1 << 40
update_s = 0.0
n_updates = 0


def apply_update(b):
    """rank 0 decides and applies; every rank receives the new blob of the touched strip"""
    global next_pk, update_s, n_updates
    s = b % S
    t0 = time.perf_counter()
    blob = None
    if rank == 0:
        cde = cdes[s]
        src_b = b - 1 - args.lookahead  # the sub-batch whose verdicts steer this update
        if src_b < 0:
            update_s += time.perf_counter() - t0
            if world > 1:  # nothing to decide yet, but every rank takes part in the broadcast
                bcs[s].broadcast(root=0)
            return
        if args.lookahead:
            cde.wait_fence(fences.pop((src_b, s)))  # that sub-batch is done; the one in flight keeps running
        else:
            cde.sync()  # the previous sub-batch's verdicts for this strip are final
        assert keep or not args.lookahead, "--lookahead keeps every sub-batch's verdicts (rank 0 checks them)"
        prev = host_out[s][src_b if keep else src_b % 2].array
        did = False
        if placed[s] and rng.random() < 0.4:
            pk = placed[s].pop(int(rng.integers(len(placed[s]))))
            cde.remove_item(pk)
            updates.append((b, s, "remove", -1, None, pk))
            did = True
        else:
            free = np.nonzero(prev == 0)[0]
            if len(free):
                q = int(free[int(rng.integers(len(free)))])
                item = int(np.searchsorted(offs, q, side="right") - 1)
                t4 = host_t[s][src_b % N_SETS].array[q].copy()
                pk = next_pk
                next_pk += 1
                cde.place_item(item, shapes[item], t4, pk)
                placed[s].append(pk)
                updates.append((b, s, "place", item, t4, pk))
                did = True
        if did:
            n_updates += 1
        if world == 1:
            cde.sync_layout()
    if world > 1:
        # collective: rank 0 flattens + uploads its own copy stream-ordered and sends; the others bind the received blob
        bcs[s].broadcast(root=0)
    update_s += time.perf_counter() - t0


# warm-up (no updates), then the timed sweep
for b in range(3):
    run_sub(b, b % 2)
barrier()
launches0 = sum(c.stats()["launches"] for c in cdes)
t0 = time.perf_counter()
for b in range(n_sub):
    if b > 0:
        apply_update(b)
    run_sub(b, b if keep else b % 2)
barrier()
dt = time.perf_counter() - t0
launches = sum(c.stats()["launches"] for c in cdes) - launches0
update_s_max = update_s
if dist is not None:
    tt = torch.tensor([dt, update_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt, update_s_max = float(tt[0]), float(tt[1])

if rank == 0:
    total = per_strip * S * n_sub * world
    line = {"metric": "collision queries/sec (MSPP sweep with layout updates, detect)", "value": total / dt, "unit": "queries/s",
            "n_gpus": world, "scaling": "weak", "seconds": dt, "queries_total": total, "sub_batches_per_gpu": n_sub,
            "queries_per_sub_batch": per_strip * S, "gpu_launches": int(launches),
            "config": {"workload": f"{args.instance}: {S} strips with minimum separation (workloads/{args.instance}_ms*.npz), "
                                   f"{n_items} item types, host buffers (16 B in, 1 B out per query), one layout update + blob "
                                   "broadcast before every sub-batch", "hazards_per_strip": [len(w.hazards) for w in strips]},
            "lookahead": args.lookahead, "layout_updates": n_updates, "update_seconds_rank0": update_s,
            "layout_broadcast": bcs[0].info if bcs else None,
            "update_ms_each": update_s / max(n_sub - 1, 1) * 1e3,
            "update_ms_each_slowest_rank": update_s_max / max(n_sub - 1, 1) * 1e3}
    if keep:
        for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            sys.path.insert(0, p)
        import orc
        from helpers import oracle_world
        orc.build()
        worlds = [oracle_world(orc, wl) for wl in strips]
        pkmap = {}
        mism = checked = 0
        cores = os.cpu_count() or 1
        ui = 0
        t_cpu = 0.0
        for b in range(n_sub):
            while ui < len(updates) and updates[ui][0] == b:
                _, s, op, item, t4, pk = updates[ui]
                ui += 1
                if op == "place":
                    # the oracle composes (rot, tx, ty) itself: recover the rotation from the item's allowed set
                    rots = strips[s].items[item].rotations
                    cands = [0.0] if rots is not None and len(rots) == 0 else list(rots)
                    rot = next(r for r in cands if np.array_equal(jg.compose(np.array([[r, t4[2], t4[3]]], np.float32))[0], t4))
                    pkmap[pk] = worlds[s].place_item(item, float(rot), float(t4[2]), float(t4[3]))
                else:
                    worlds[s].remove_item(pkmap.pop(pk))
            for s in range(S):
                tq = host_t[s][b % N_SETS].array
                tc = time.perf_counter()
                exp = np.concatenate([worlds[s].detect_batch(i, np.ascontiguousarray(tq[offs[i]:offs[i + 1]]), nthreads=cores)
                                      for i in range(n_items)])
                t_cpu += time.perf_counter() - tc
                mism += int((exp != host_out[s][b].array).sum())
                checked += len(exp)
        line["checked_queries"] = checked
        line["mismatches_vs_oracle"] = mism
        line["cpu_baseline"] = {"value": checked / t_cpu, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"rank 0's {checked} queries, oracle/ C++ restatement, layout replayed"}
    print(json.dumps(line))

for c in cdes:
    c.close()
if dist is not None:
    dist.destroy_process_group()
