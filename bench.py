#!/usr/bin/env python
"""bench.py — collision queries/s of the batched CDE on B200 (BASELINE.json config 2).

A step = one pass of the hot path (detect_surrogate_collision + detect_poly_collision) over one
1,048,576-query batch of the swim instance against the half-filled swim layout: 10 item types x
~104,858 seeded queries each (workloads/swim_half.npz, tools/make_workloads.py).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scaling weak|strong]

value      queries/s, inputs resident in HBM, CUDA events on the engine's compute stream (the stream the kernels are
           launched on), L2 flushed (untimed) before every step.
e2e        queries/s through the public host-buffer API on pinned host arrays, every step's H2D and D2H inside the
           timed region, K batches enqueued back to back and one sync at the end. The headline form is the one a
           caller whose items have discrete allowed orientations uses (swim: {0, pi}): jg_detect_multi_indexed, rotation
           index u8 + translation 2 x f32 = 9 B/query in, 1 B/query out. Beside it: the composed 16 B form
           (e2e.composed_input), the 5 B lattice form (e2e.lattice_input), and each form one batch at a time.
strong     ONE 1M-query host batch split over the N GPUs by the in-process pool (jg_pool_*), results gathered into one
           pinned array, timed on rank 0 (wall clock: the caller's view) while the other ranks idle.
roofline   FP32 pipe without FMA (the kernels are compiled -fmad=false for bit-exactness): algorithmic
           flops counted by the instrumented oracle (workloads/swim_half.flops.json) / device time,
           against the peak measured live by jg_fp32_nofma_peak. The HBM view of the same step is
           reported beside it (17 algorithmic bytes per query).
cpu_baseline  the CPU restatement of the reference (oracle/, -O3 -march=native, persistent thread pool over all host
           threads, dynamic chunks) on the same batch, and a full-batch verdict comparison (mismatches must be 0).
N > 1      one process per GPU (torchrun); the layout blob is flattened on rank 0 and broadcast over NCCL inside the
           library (jg_layout_broadcast), every rank resolves its own 1M-query batch (weak scaling), no data-path
           collective. --scaling strong: ONE batch, sharded [r n/N, (r+1) n/N) over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT,):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "collision queries/sec (swim, 1M placements, detect)"
UNIT = "queries/s"
WORKLOAD = "swim_half"
TOTAL_QUERIES = 1 << 20
ALG_BYTES_PER_QUERY = 17  # 16 B (cos, sin, tx, ty) in + 1 B verdict out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--queries", type=int, default=TOTAL_QUERIES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the in-process pool leg")
    return ap.parse_args()


def make_config(args, flops_info=None) -> dict:
    """The `config` object of the JSON line: identical in both arms (--impl b200 / reference)."""
    per = "per GPU" if args.scaling == "weak" else "in total, sharded over the GPUs"
    return {"workload": f"{args.workload}: {args.queries} uniform queries {per} over the CDE bbox, 10 swim item types, "
                        "24 placed items + exterior, quadtree depth 5, cd_threshold 16, 2 ff poles",
            "queries_per_step": args.queries, "scaling": args.scaling,
            "l2": "flushed (256 MiB fill, untimed) before every timed step of `value`",
            "timing": "`value`: CUDA events on the launching stream around each step; the step's launches are enqueued behind a gate kernel (jg_timer_start_gated), so the interval is device time and the host's enqueue time is not in it",
            "verdict_split": (flops_info or {}).get("verdicts")}


def load_flops(args):
    try:
        with open(os.path.join(ROOT, "workloads", f"{args.workload}.flops.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def load_oracle():
    """CPU restatement of the reference (test infrastructure): only the cpu_baseline / --impl reference
    legs come here."""
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import orc
    from helpers import oracle_world
    orc.build()
    return orc, oracle_world


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU implementation of the path on the host cores. The Rust crate cannot be
    built in this image (no cargo), so this is the oracle port: one parallel pass per step over the whole segmented batch,
    all host threads. Nothing of the product is loaded here: the workload fixture is read with numpy and the queries are
    composed by the oracle's own compose (the same libm call)."""
    if rank != 0:
        return
    orc, oracle_world = load_oracle()
    from importlib import util as _u
    spec = _u.spec_from_file_location("_jg_workload", os.path.join(ROOT, "jagua-rs_b200", "workload.py"))
    wlmod = _u.module_from_spec(spec)  # (a plain-numpy module; loading it by path keeps libjagua_cuda.so out of this process)
    sys.modules["_jg_workload"] = wlmod
    spec.loader.exec_module(wlmod)
    wl = wlmod.load_workload(os.path.join(ROOT, "workloads", f"{args.workload}.npz"))
    w = oracle_world(orc, wl)
    sizes = wlmod.split_batch(args.queries, len(wl.items))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ids = list(range(len(wl.items)))
    t = np.ascontiguousarray(np.concatenate([wlmod.make_queries(wl, i, n, seed=0, compose_fn=orc.compose) for i, n in enumerate(sizes)]))
    out = np.zeros(len(t), np.uint8)
    cores = os.cpu_count() or 1

    def step():
        w.detect_multi(ids, offs, t, nthreads=cores, out=out)

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.queries * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, load_flops(args)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"the full {args.queries}-query batch per step; oracle/ C++ restatement built -O3 -march=native "
                                       "-ffp-contract=off, persistent thread pool, dynamic chunks of 2048 queries"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE line (the JSON record): libraries that write to fd 1 while this runs - NCCL prints its version
    # banner there when the first communicator is created - are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import jagua_rs_b200 as jg
    from jagua_rs_b200 import multigpu

    # pinned staging buffers on the GPU's own socket (matters from 4 ranks up: every rank streams ~53 GB/s)
    # (N = 1 keeps every core: the cpu_baseline leg below uses all of them)
    numa = multigpu.bind_to_gpu_numa_node(local_rank) if world > 1 and not os.environ.get("JG_NO_NUMA_BIND") else {"why": "single rank or JG_NO_NUMA_BIND"}

    dist = None
    host_group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # host-side barrier for the legs in which only rank 0 drives the GPUs

    wl = jg.load_workload(args.workload)
    flops_info = load_flops(args)

    # ---- layout: built + flattened on rank 0, broadcast over NCCL inside the library, bound on every rank
    t_b0 = time.perf_counter()
    bcast = None
    if world > 1:
        if rank == 0:
            cde = jg.build_engine(wl, device=local_rank)
        else:
            cde = jg.CDEngine(wl.cde_bbox, wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, device=local_rank)
            for i, it in enumerate(wl.items):
                cde.upload_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers))
        bcast = multigpu.LayoutBroadcaster(cde, dist, rank, world)
        bcast.broadcast(root=0)
    else:
        cde = jg.build_engine(wl, device=local_rank)
        cde.sync_layout()
    layout_setup_ms = (time.perf_counter() - t_b0) * 1e3

    # ---- inputs. weak: every rank its own seeded 1M batch; strong: ONE batch (seed 0), rank r owns the shard [r n/N, (r+1) n/N)
    n_items = len(wl.items)
    sizes = jg.split_batch(args.queries, n_items)
    full_offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    seed = rank if args.scaling == "weak" else 0
    full_t = np.concatenate([jg.make_queries(wl, i, n, seed=seed) for i, n in enumerate(sizes)]).astype(np.float32)
    indexed_ok = all(it.rotations is not None and len(it.rotations) > 0 for it in wl.items)
    full_ridx = full_txy = full_ij = lattice = None
    if indexed_ok:
        parts = [jg.make_queries_indexed(wl, i, n, seed=seed) for i, n in enumerate(sizes)]
        full_ridx = np.concatenate([p[0] for p in parts])
        full_txy = np.concatenate([p[1] for p in parts])
        lparts = [jg.make_queries_lattice(wl, i, n, seed=seed) for i, n in enumerate(sizes)]
        lattice = lparts[0][0]
        full_lridx = np.concatenate([p[1] for p in lparts])
        full_ij = np.concatenate([p[2] for p in lparts])
        for i in range(n_items):
            cde.set_item_rotations(i, wl.items[i].rotations)
    if args.scaling == "strong" and world > 1:
        lo, hi, seg_ids, offs = multigpu.shard_segments(full_offs, rank, world)
        item_ids = seg_ids
    else:
        lo, hi, item_ids, offs = 0, int(full_offs[-1]), list(range(n_items)), full_offs
    n_q = hi - lo
    host_t = jg.PinnedBuffer((n_q, 4), np.float32)
    host_t.array[:] = full_t[lo:hi]
    host_out = jg.PinnedBuffer((n_q,), np.uint8)
    dev_t = jg.DeviceBuffer(cde, n_q * 16)
    dev_t.upload(host_t.array)
    dev_out = jg.DeviceBuffer(cde, n_q)

    # one step = ONE multi-item call: the 10 item segments of the batch go through one launch sequence
    def step_device():
        cde.detect_multi_device(item_ids, offs, dev_t.ptr, dev_out.ptr)

    host_ridx = jg.PinnedBuffer((n_q,), np.uint8)
    host_txy = jg.PinnedBuffer((n_q, 2), np.float32)
    host_lridx = jg.PinnedBuffer((n_q,), np.uint8)
    host_ij = jg.PinnedBuffer((n_q, 2), np.uint16)
    if indexed_ok:
        host_ridx.array[:] = full_ridx[lo:hi]
        host_txy.array[:] = full_txy[lo:hi]
        host_lridx.array[:] = full_lridx[lo:hi]
        host_ij.array[:] = full_ij[lo:hi]

    def barrier():
        cde.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    peak_gflops, _ = cde.fp32_nofma_peak()

    # ---- value: device-resident, CUDA events, L2 flushed (untimed) before every step
    for _ in range(max(args.warmup, 3)):
        step_device()
    cde.sync()
    launches0 = cde.stats()["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, step_ms = 0.0, []
    for _ in range(args.steps):
        cde.flush_l2()
        cde.timer_start(gated=True)
        step_device()
        step_ms.append(cde.timer_stop())
        dev_ms += step_ms[-1]
    barrier()
    gpu_launches = cde.stats()["launches"] - launches0 - args.steps  # minus the untimed L2 flush fills

    # ---- e2e: host buffers through the public API, copies inside the timed region
    # (a) one batch at a time: enqueue + sync per step (the library cuts an idle engine's batch into pieces of 256 Ki queries)
    # (b) batches in flight: K steps enqueued back to back (whole batch per piece stream), one sync at the end
    ring = {k: [jg.PinnedBuffer((n_q,), np.uint8) for _ in range(4)] for k in ("composed", "indexed", "lattice")}
    ring["composed"][0] = host_out

    def run_e2e(step_fn, pipelined):
        cde.set_batching(0, 4)  # the library's per-batch piece schedule, 4 piece streams
        for i in range(max(args.warmup, 3)):
            step_fn(i)
            cde.sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_fn(i)
            if not pipelined:
                cde.sync()
        barrier()
        return time.perf_counter() - t0

    forms = {"composed": lambda i: cde.detect_multi(item_ids, offs, host_t.array, out=ring["composed"][i % 4].array, wait=False)}
    if indexed_ok:
        forms["indexed"] = lambda i: cde.detect_multi_indexed(item_ids, offs, host_ridx.array, host_txy.array, out=ring["indexed"][i % 4].array, wait=False)
        forms["lattice"] = lambda i: cde.detect_multi_lattice(item_ids, offs, lattice, host_lridx.array, host_ij.array, out=ring["lattice"][i % 4].array, wait=False)
    names = list(forms)
    e2e_s = {k: [run_e2e(forms[k], False), run_e2e(forms[k], True)] for k in names}  # [serial, pipelined]
    clocks = sampler.stop()

    if dist is not None:
        import torch
        flat = [dev_ms] + [x for k in names for x in e2e_s[k]]
        tt = torch.tensor(flat, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in tt]
        dev_ms = vals[0]
        for j, k in enumerate(names):
            e2e_s[k] = vals[1 + 2 * j: 3 + 2 * j]

    # verdicts of the e2e steps vs the device-resident path (same inputs -> identical); the lattice batch holds other
    # (quantised) placements: it is checked against the composed records it stands for, through the device path
    dev_verdicts = dev_out.download((n_q,), np.uint8)
    n_ring = min(4, args.steps)  # every in-flight result buffer the timed steps wrote
    same = {k: all(np.array_equal(ring[k][j].array, dev_verdicts) for j in range(n_ring)) for k in names if k != "lattice"}
    if "lattice" in names:
        lt = np.ascontiguousarray(np.concatenate([jg.compose_lattice_queries(wl, item_ids[s], lattice, host_lridx.array[offs[s]:offs[s + 1]],
                                                                             host_ij.array[offs[s]:offs[s + 1]]) for s in range(len(item_ids))]))
        lat_exp = cde.detect_multi(item_ids, offs, lt)
        same["lattice"] = all(np.array_equal(ring["lattice"][j].array, lat_exp) for j in range(n_ring))

    # ---- strong: ONE 1M-query host batch over all N GPUs through the in-process pool, on rank 0 (the other ranks idle)
    strong = None
    if not args.no_strong and indexed_ok:
        if host_group is not None:
            dist.barrier(group=host_group)
        if rank == 0:
            try:  # (a failure here must not leave the other ranks waiting at the barrier below)
                strong = run_strong(jg, wl, args, world, full_offs, full_t, full_ridx, full_txy)
            except Exception as ex:  # noqa: BLE001 - reported in the line
                strong = {"error": repr(ex)}
        if host_group is not None:
            dist.barrier(group=host_group)

    if rank == 0:
        total_q = args.queries * world if args.scaling == "weak" else args.queries
        value = total_q * args.steps / (dev_ms * 1e-3)
        ms_per_step = dev_ms / args.steps
        per_gpu_q = n_q
        flops_per_q = flops_info.get("detect_flops_per_query") if (args.queries == flops_info.get("queries") and args.scaling == "weak") else None
        roofline = None
        traffic = None
        issue = None
        clocks_mhz_guess = clocks.get("sm_mhz") or 1965.0
        try:  # DRAM bytes and warp instructions of the two kernels: STATIC numbers from the committed ncu capture of this step
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            if cde.detect_path() == "grid" and per_gpu_q == TOTAL_QUERIES:
                traffic = {"bytes_per_step": tr["k_s1_grid"] + tr["k_s2_grid"] + tr.get("k_s2_contain", 0), "algorithmic_bytes_per_step": ALG_BYTES_PER_QUERY * args.queries,
                           "measured": "static: read from profiles/ncu_traffic.json, not measured in this run", "source": tr["source"]}
                wi = tr.get("warp_instructions")
                if wi:  # what actually bounds these kernels: warp instructions issued vs the SMs' issue slots
                    n_wi = wi["k_s1_grid"] + wi["k_s2_grid"] + wi.get("k_s2_contain", 0)
                    peak_issue = 148 * 4 * clocks_mhz_guess * 1e6
                    issue = {"warp_instructions_per_step": n_wi, "achieved": n_wi / (ms_per_step * 1e-3) / 1e9,
                             "peak": peak_issue / 1e9, "unit": "G warp-instructions/s", "frac": n_wi / (ms_per_step * 1e-3) / peak_issue,
                             "peak_source": "148 SMs x 4 schedulers x 1 instruction/clk x the SM clock sampled during the run",
                             "measured": "static instruction count (profiles/ncu_traffic.json) / this run's device time",
                             "source": wi["source"]}
        except (OSError, KeyError, ValueError):
            pass
        if flops_per_q:
            ach = flops_per_q * args.queries / (ms_per_step * 1e-3) / 1e12  # per GPU
            roofline = {"bound": "fp32_nofma", "achieved": ach, "peak": peak_gflops / 1e3, "unit": "TFLOP/s",
                        "frac": ach / (peak_gflops / 1e3), "traffic": traffic["bytes_per_step"] if traffic else None,
                        "traffic_detail": traffic,
                        "peak_source": "measured live: jg_fp32_nofma_peak (independent FADD/FMUL chains)",
                        "issue": issue, "flops_per_query": flops_per_q, "kernels": f"{cde.detect_path()} path: " + ("k_s1_grid + k_s2_grid + k_s2_contain (containment tests in their own kernel from 640 Ki queries per launch)" if cde.detect_path() == "grid" else "k_s1_surr + k_s2_poly") + ", one launch sequence for the 10 item segments",
                        "hbm": {"achieved": ALG_BYTES_PER_QUERY * args.queries / (ms_per_step * 1e-3) / 1e9, "peak": 6554.2,
                                "unit": "GB/s", "frac": ALG_BYTES_PER_QUERY * args.queries / (ms_per_step * 1e-3) / 1e9 / 6554.2,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs"}}

        def e2e_obj(k, bytes_in, desc):
            ser, pip = e2e_s[k]
            return {"value": total_q * args.steps / pip, "unit": UNIT, "ms_per_step": pip / args.steps * 1e3,
                    "h2d_bytes_per_step": bytes_in * n_q, "d2h_bytes_per_step": n_q, "input": desc, "same_verdicts_as_device_path": bool(same[k]),
                    "one_batch_at_a_time": {"value": total_q * args.steps / ser, "unit": UNIT, "ms_per_step": ser / args.steps * 1e3,
                                            "mode": "sync after every batch; the library's piece schedule for an idle engine: pieces of 256 Ki queries on 4 streams (jg_engine_set_batching(0, 4))"}}

        composed = e2e_obj("composed", 16, "jg_detect_multi: composed (cos, sin, tx, ty) f32 records, 16 B/query, pinned host memory")
        head = composed
        if indexed_ok:
            head = e2e_obj("indexed", 9, "jg_detect_multi_indexed: rotation index u8 + translation 2 x f32 = 9 B/query - the form for items with a "
                                         "discrete set of allowed orientations (RotationRange::Discrete; swim: {0, pi}); same queries, pinned host memory")
            head["composed_input"] = composed
            head["lattice_input"] = e2e_obj("lattice", 5, "jg_detect_multi_lattice: rotation index u8 + lattice coordinates 2 x u16 = 5 B/query (the same uniform "
                                                          "distribution quantised to 65536 steps per axis; verdicts checked against the composed records)")
        head["mode"] = ("K steps enqueued back to back (4 rotating result buffers), one sync at the end: consecutive batches overlap on the "
                        "engine's streams; every step's H2D and D2H is inside the timed region")
        head["pcie_h2d_ceiling_queries_per_s"] = 53e9 / (9 if indexed_ok else 16)
        sm = np.array(step_ms)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": make_config(args, flops_info),
                "clocks": clocks, "gpu_launches": int(gpu_launches),
                "step_ms_spread": {"min": float(sm.min()), "median": float(np.median(sm)), "max": float(sm.max()), "n": len(sm), "rank": 0},
                "e2e": head, "strong": strong,
                "roofline": roofline, "layout_setup_ms": layout_setup_ms, "numa_binding_rank0": numa,
                "layout_broadcast": bcast.info if bcast is not None else None}
        if world == 1 and not args.no_cpu_baseline:
            orc, oracle_world = load_oracle()
            w = oracle_world(orc, wl)
            cores = os.cpu_count() or 1
            qs = np.ascontiguousarray(host_t.array)
            exp = w.detect_multi(item_ids, offs, qs, nthreads=cores)  # warm-up + checker
            mism = int((exp != host_out.array).sum())
            if indexed_ok:
                mism += int((exp != ring["indexed"][0].array).sum())
            reps, t0 = 0, time.perf_counter()
            while reps < 60 and time.perf_counter() - t0 < 10.0:
                w.detect_multi(item_ids, offs, qs, nthreads=cores, out=exp)
                reps += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": args.queries * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{reps} passes over the full {args.queries}-query batch ({dt:.1f} s), oracle/ C++ "
                                              "restatement built -O3 -march=native -ffp-contract=off, persistent thread pool, dynamic chunks",
                                    "mismatches_vs_gpu": mism}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    cde.close()
    if dist is not None:
        dist.destroy_process_group()


def run_strong(jg, wl, args, n_dev, offs, t, ridx, txy):
    """ONE host batch over n_dev GPUs of this process: jg_pool_* shards it, every device writes its slice of one pinned
    result array. Wall clock around the caller's calls (this is the caller's view: enqueue, sync, results on the host)."""
    host = jg.build_engine(wl, device=None)
    pool = jg.EnginePool(host, list(range(n_dev)))
    try:
        pool.upload_items(wl.items)
        for i in range(len(wl.items)):
            pool.set_item_rotations(i, wl.items[i].rotations)
        pool.sync_layout()
        n = int(offs[-1])
        ids = list(range(len(wl.items)))
        p_r, p_t = jg.PinnedBuffer((n,), np.uint8), jg.PinnedBuffer((n, 2), np.float32)
        p_r.array[:] = ridx
        p_t.array[:] = txy
        p_c = jg.PinnedBuffer((n, 4), np.float32)
        p_c.array[:] = t
        outs = [jg.PinnedBuffer((n,), np.uint8) for _ in range(4)]
        res = {}
        for form in ("indexed", "composed"):
            for pipelined in (False, True):
                per_dev = (n + n_dev - 1) // n_dev
                # every engine cuts its shard by the library's per-batch schedule (detect_multi_host); one at a time = the blocking
                # call (enqueue + wait on the device's own thread)
                pool.set_batching(0, 4)

                def step(i):
                    if form == "indexed":
                        pool.detect_multi_indexed(ids, offs, p_r.array, p_t.array, out=outs[i % 4].array, wait=not pipelined)
                    else:
                        pool.detect_multi(ids, offs, p_c.array, out=outs[i % 4].array, wait=not pipelined)

                for i in range(max(args.warmup, 3)):
                    step(i)
                    pool.sync()
                t0 = time.perf_counter()
                for i in range(args.steps):
                    step(i)
                pool.sync()
                dt = time.perf_counter() - t0
                res[(form, pipelined)] = dt
        ref = outs[0].array.copy()
        ok = all(np.array_equal(o.array, ref) for o in outs[:min(4, args.steps)])
        single = jg.build_engine(wl, device=0)
        exp = single.detect_multi(ids, offs, np.ascontiguousarray(t))
        single.close()
        ok = ok and np.array_equal(ref, exp)

        def obj(form, b):
            return {"value": n * args.steps / res[(form, True)], "unit": UNIT, "ms_per_batch": res[(form, True)] / args.steps * 1e3,
                    "one_batch_at_a_time": {"value": n * args.steps / res[(form, False)], "unit": UNIT,
                                            "ms_per_batch": res[(form, False)] / args.steps * 1e3},
                    "h2d_bytes_per_batch": b * n, "d2h_bytes_per_batch": n}

        out = obj("indexed", 9)
        out["composed_input"] = obj("composed", 16)
        out.update({"n_gpus": n_dev, "queries_per_batch": n, "scaling": "strong", "same_verdicts_as_single_engine": bool(ok),
                    "mode": "jg_pool_detect_multi_indexed_async: one 9 B/query host batch, contiguous shards, per-device D2H into disjoint "
                            "slices of one pinned result array; wall clock on the calling thread"})
        return out
    finally:
        pool.close()
        host.close()


if __name__ == "__main__":
    main()
