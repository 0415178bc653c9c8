#!/usr/bin/env python
"""bench.py — collision queries/s of the batched CDE on B200 (BASELINE.json config 2).

A step = one pass of the hot path (detect_surrogate_collision + detect_poly_collision) over one
1,048,576-query batch of the swim instance against the half-filled swim layout: 10 item types x
~104,858 seeded (cos, sin, tx, ty) queries each (workloads/swim_half.npz, tools/make_workloads.py).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

value      queries/s, inputs resident in HBM, CUDA events on the engine's compute stream, L2 flushed
           (untimed) before every step.
e2e        queries/s through the public host-buffer API (CDEngine.detect_multi on pinned host arrays, the
           asynchronous form: K batches enqueued back to back, one sync at the end, so consecutive batches
           overlap their copies and kernels on the engine's streams): H2D of every step's transforms and
           D2H of every step's verdicts inside the timed region. e2e.one_batch_at_a_time is the same call
           followed by a sync after every batch (the latency-bound figure).
roofline   FP32 pipe without FMA (the kernels are compiled -fmad=false for bit-exactness): algorithmic
           flops counted by the instrumented oracle (workloads/swim_half.flops.json) / device time,
           against the peak measured live by jg_fp32_nofma_peak. The HBM view of the same step is
           reported beside it (17 algorithmic bytes per query).
cpu_baseline  the CPU restatement of the reference (oracle/, all host threads) on the same batch, and
           a full-batch verdict comparison (mismatches must be 0).
N > 1      one process per GPU (torchrun); the layout blob is flattened on rank 0 and broadcast over
           NCCL, every rank resolves its own 1M-query batch (weak scaling), no data-path collective.
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
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--queries", type=int, default=TOTAL_QUERIES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


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
    """--impl reference: the reference's CPU implementation of the path on the host cores. The Rust
    crate cannot be built in this image (no cargo), so this is the oracle port, all host threads."""
    if rank != 0:
        return
    import jagua_rs_b200 as jg
    orc, oracle_world = load_oracle()
    wl = jg.load_workload(args.workload)
    w = oracle_world(orc, wl)
    sizes = jg.split_batch(args.queries, len(wl.items))
    qs = [jg.make_queries(wl, i, n, seed=0) for i, n in enumerate(sizes)]
    cores = os.cpu_count() or 1

    def step():
        for i, q in enumerate(qs):
            w.detect_batch(i, q, nthreads=cores)

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.queries * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {args.queries} uniform queries over the CDE bbox, 10 swim item types, "
                                   "24 placed items + exterior, quadtree depth 5", "queries_per_step": args.queries},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"the full {args.queries}-query batch per step, std::thread static sharding"},
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

    import jagua_rs_b200 as jg
    from jagua_rs_b200 import multigpu

    # pinned staging buffers on the GPU's own socket (matters from 4 ranks up: every rank streams ~53 GB/s)
    # (N = 1 keeps every core: the cpu_baseline leg below uses all of them)
    numa = multigpu.bind_to_gpu_numa_node(local_rank) if world > 1 and not os.environ.get("JG_NO_NUMA_BIND") else {"why": "single rank or JG_NO_NUMA_BIND"}

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = jg.load_workload(args.workload)
    with open(os.path.join(ROOT, "workloads", f"{args.workload}.flops.json")) as f:
        flops_info = json.load(f)

    # ---- layout: built + flattened on rank 0, broadcast (NCCL) and uploaded from device memory elsewhere
    t_b0 = time.perf_counter()
    if world > 1:
        import torch
        from jagua_rs_b200 import multigpu
        blob = None
        if rank == 0:
            cde = jg.build_engine(wl, device=local_rank)
            blob = cde.flatten()
        else:
            cde = jg.CDEngine(wl.cde_bbox, wl.quadtree_depth, wl.cd_threshold, wl.n_ff_poles, wl.n_ff_piers, device=local_rank)
            for i, it in enumerate(wl.items):
                cde.upload_item(i, jg.ItemShape(it.verts, it.poi, it.poles, it.piers))
        d_blob = multigpu.broadcast_blob(blob, dist, src=0, device=f"cuda:{local_rank}")
        torch.cuda.synchronize()
        cde.upload_layout_device(d_blob.data_ptr(), d_blob.numel(), epoch=1)
    else:
        cde = jg.build_engine(wl, device=local_rank)
        cde.sync_layout()
    layout_setup_ms = (time.perf_counter() - t_b0) * 1e3

    # ---- inputs: every rank its own seeded batch (weak scaling); pinned host + device-resident copies
    sizes = jg.split_batch(args.queries, len(wl.items))
    n_items = len(wl.items)
    item_ids = list(range(n_items))
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n_q = int(offs[-1])
    host_t = jg.PinnedBuffer((n_q, 4), np.float32)
    for i, n in enumerate(sizes):
        host_t.array[offs[i]:offs[i + 1]] = jg.make_queries(wl, i, n, seed=rank)
    host_out = jg.PinnedBuffer((n_q,), np.uint8)
    dev_t = jg.DeviceBuffer(cde, n_q * 16)
    dev_t.upload(host_t.array)
    dev_out = jg.DeviceBuffer(cde, n_q)

    # one step = ONE multi-item call: the 10 item segments of the batch go through one launch sequence
    def step_device():
        cde.detect_multi_device(item_ids, offs, dev_t.ptr, dev_out.ptr)

    # the same batch in the compact wire form (rotation index u8 + translation 2 x f32 = 9 B/query)
    host_ridx = jg.PinnedBuffer((n_q,), np.uint8)
    host_txy = jg.PinnedBuffer((n_q, 2), np.float32)
    host_out_ix = jg.PinnedBuffer((n_q,), np.uint8)
    indexed_ok = all(it.rotations is not None and len(it.rotations) > 0 for it in wl.items)
    if indexed_ok:
        for i, n in enumerate(sizes):
            cde.set_item_rotations(i, wl.items[i].rotations)
            r, t = jg.make_queries_indexed(wl, i, n, seed=rank)
            host_ridx.array[offs[i]:offs[i + 1]] = r
            host_txy.array[offs[i]:offs[i + 1]] = t

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
    dev_ms = 0.0
    for _ in range(args.steps):
        cde.flush_l2()
        cde.timer_start()
        step_device()
        dev_ms += cde.timer_stop()
    barrier()
    gpu_launches = cde.stats()["launches"] - launches0 - args.steps  # minus the untimed L2 flush fills

    # ---- e2e: host buffers through the public API, copies inside the timed region
    # (a) one batch at a time: enqueue + sync per step, pieces of 256 Ki queries on 4 streams
    # (b) batches in flight: K steps enqueued back to back (whole batch per stream, 3 streams), one sync at the end
    ring = [host_out] + [jg.PinnedBuffer((n_q,), np.uint8) for _ in range(3)]
    ring_ix = [host_out_ix] + [jg.PinnedBuffer((n_q,), np.uint8) for _ in range(3)]

    def run_e2e(step_fn, pipelined):
        cde.set_batching(1 << 20 if pipelined else 1 << 18, 3 if pipelined else 4)
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

    def enq(i):
        cde.detect_multi(item_ids, offs, host_t.array, out=ring[i % 4].array, wait=False)

    def enq_ix(i):
        cde.detect_multi_indexed(item_ids, offs, host_ridx.array, host_txy.array, out=ring_ix[i % 4].array, wait=False)

    e2e_serial_s = run_e2e(enq, False)
    e2e_s = run_e2e(enq, True)
    e2e_ix_s = e2e_ix_serial_s = 0.0
    if indexed_ok:
        e2e_ix_serial_s = run_e2e(enq_ix, False)
        e2e_ix_s = run_e2e(enq_ix, True)
    clocks = sampler.stop()

    if dist is not None:
        import torch
        tt = torch.tensor([dev_ms, e2e_s, e2e_ix_s, e2e_serial_s, e2e_ix_serial_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_ix_s, e2e_serial_s, e2e_ix_serial_s = (float(x) for x in tt)

    # verdicts of the last e2e step vs the device-resident path (same inputs -> identical)
    dev_verdicts = dev_out.download((n_q,), np.uint8)
    n_ring = min(4, args.steps)  # every in-flight result buffer the timed steps wrote
    same = all(np.array_equal(ring[k].array, dev_verdicts) for k in range(n_ring))
    same_ix = bool(indexed_ok and all(np.array_equal(ring_ix[k].array, dev_verdicts) for k in range(n_ring)))

    if rank == 0:
        total_q = args.queries * world
        value = total_q * args.steps / (dev_ms * 1e-3)
        e2e = total_q * args.steps / e2e_s
        ms_per_step = dev_ms / args.steps
        flops_per_q = flops_info["detect_flops_per_query"] if args.queries == flops_info["queries"] else None
        roofline = None
        traffic = None
        issue = None
        clocks_mhz_guess = clocks.get("sm_mhz") or 1965.0
        try:  # DRAM bytes of the two kernels from the committed ncu capture of this same step
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            if cde.detect_path() == "grid" and args.queries == TOTAL_QUERIES:
                traffic = {"bytes_per_step": tr["k_s1_grid"] + tr["k_s2_grid"], "algorithmic_bytes_per_step": ALG_BYTES_PER_QUERY * args.queries,
                           "source": tr["source"]}
                wi = tr.get("warp_instructions")
                if wi:  # what actually bounds these kernels: warp instructions issued vs the SMs' issue slots
                    n_wi = wi["k_s1_grid"] + wi["k_s2_grid"]
                    peak_issue = 148 * 4 * clocks_mhz_guess * 1e6
                    issue = {"warp_instructions_per_step": n_wi, "achieved": n_wi / (ms_per_step * 1e-3) / 1e9,
                             "peak": peak_issue / 1e9, "unit": "G warp-instructions/s", "frac": n_wi / (ms_per_step * 1e-3) / peak_issue,
                             "peak_source": "148 SMs x 4 schedulers x 1 instruction/clk x the SM clock sampled during the run",
                             "source": wi["source"]}
        except (OSError, KeyError, ValueError):
            pass
        if flops_per_q:
            ach = flops_per_q * args.queries / (ms_per_step * 1e-3) / 1e12  # per GPU
            roofline = {"bound": "fp32_nofma", "achieved": ach, "peak": peak_gflops / 1e3, "unit": "TFLOP/s",
                        "frac": ach / (peak_gflops / 1e3), "traffic": traffic["bytes_per_step"] if traffic else None,
                        "traffic_detail": traffic,
                        "peak_source": "measured live: jg_fp32_nofma_peak (independent FADD/FMUL chains)",
                        "issue": issue, "flops_per_query": flops_per_q, "kernels": f"{cde.detect_path()} path: " + ("k_s1_grid + k_s2_grid" if cde.detect_path() == "grid" else "k_s1_surr + k_s2_poly") + ", one launch pair for the 10 item segments",
                        "hbm": {"achieved": ALG_BYTES_PER_QUERY * args.queries / (ms_per_step * 1e-3) / 1e9, "peak": 6555.2,
                                "unit": "GB/s", "frac": ALG_BYTES_PER_QUERY * args.queries / (ms_per_step * 1e-3) / 1e9 / 6555.2,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: {args.queries} uniform queries per GPU over the CDE bbox, 10 swim item types, "
                                       "24 placed items + exterior, quadtree depth 5, cd_threshold 16, 2 ff poles",
                           "queries_per_step_per_gpu": args.queries, "l2": "flushed (256 MiB fill, untimed) before every timed step",
                           "verdict_split": flops_info.get("verdicts")},
                "clocks": clocks, "gpu_launches": int(gpu_launches),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 16 * args.queries, "d2h_bytes_per_step": args.queries,
                        "ms_per_step": e2e_s / args.steps * 1e3, "same_verdicts_as_device_path": bool(same),
                        "input": "composed (cos, sin, tx, ty) f32 records, 16 B/query, pinned host memory",
                        "mode": "detect_multi(wait=False) for all K steps back to back (4 rotating result buffers), one sync at the end: "
                                "consecutive batches overlap on the engine's streams; every step's H2D and D2H is inside the timed region",
                        "pcie_h2d_ceiling_queries_per_s": 53e9 / 16,
                        "one_batch_at_a_time": {"value": total_q * args.steps / e2e_serial_s, "unit": UNIT,
                                                "ms_per_step": e2e_serial_s / args.steps * 1e3,
                                                "mode": "sync after every batch; pieces of 256 Ki queries on 4 streams"},
                        "compact_input": None if not indexed_ok else {
                            "value": total_q * args.steps / e2e_ix_s, "unit": UNIT, "ms_per_step": e2e_ix_s / args.steps * 1e3,
                            "one_batch_at_a_time": {"value": total_q * args.steps / e2e_ix_serial_s, "unit": UNIT,
                                                    "ms_per_step": e2e_ix_serial_s / args.steps * 1e3},
                            "h2d_bytes_per_step": 9 * args.queries, "d2h_bytes_per_step": args.queries,
                            "input": "jg_detect_multi_indexed: rotation index u8 + translation 2 x f32 = 9 B/query (items with discrete "
                                     "allowed orientations), same queries", "same_verdicts": same_ix}},
                "roofline": roofline, "layout_setup_ms": layout_setup_ms, "numa_binding_rank0": numa}
        if world == 1 and not args.no_cpu_baseline:
            orc, oracle_world = load_oracle()
            w = oracle_world(orc, wl)
            cores = os.cpu_count() or 1
            qs = [np.ascontiguousarray(host_t.array[offs[i]:offs[i + 1]]) for i in range(n_items)]
            exp = [w.detect_batch(i, qs[i], nthreads=cores) for i in range(n_items)]  # warm-up + checker
            mism = int(sum((exp[i] != host_out.array[offs[i]:offs[i + 1]]).sum() for i in range(n_items)))
            reps, t0 = 0, time.perf_counter()
            while reps < 40 and time.perf_counter() - t0 < 10.0:
                for i in range(n_items):
                    w.detect_batch(i, qs[i], nthreads=cores)
                reps += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": args.queries * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{reps} passes over the full {args.queries}-query batch ({dt:.1f} s), oracle/ C++ "
                                              "restatement, std::thread static sharding",
                                    "mismatches_vs_gpu": mism}
        print(json.dumps(line))

    cde.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
