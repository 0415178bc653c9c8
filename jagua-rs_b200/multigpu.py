"""Multi-GPU plumbing of the query path: one process per GPU, the flattened layout is
replicated (rank 0 flattens, everyone else receives the blob), the query batch is sharded into contiguous
slices, and there is no collective on the query path. torch.distributed is only the transport
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-query batch owned by `rank` (sizes differ by at most one)."""
    assert 0 <= rank < world and n >= 0
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_segments(seg_offsets, rank: int, world: int):
    """The part of a multi-item batch (segment k = queries seg_offsets[k]..seg_offsets[k+1]) that falls into
    `rank`'s slice: (lo, hi, segment indices, offsets relative to lo) - ready for detect_multi on the slice."""
    offs = np.asarray(seg_offsets, np.int64)
    lo, hi = shard_bounds(int(offs[-1]), rank, world)
    ids, rel = [], [0]
    for k in range(len(offs) - 1):
        a, b = max(int(offs[k]), lo), min(int(offs[k + 1]), hi)
        if b > a:
            ids.append(k)
            rel.append(rel[-1] + (b - a))
    return lo, hi, ids, np.asarray(rel, np.int64)


def broadcast_blob(blob: Optional[np.ndarray], dist, src: int = 0, device: Optional[str] = None):
    """Replicates a layout blob (uint8 array, only needed on `src`) to every rank. Returns a torch uint8 tensor
    on `device` (None = CPU) holding the blob on every rank."""
    import torch

    dev = torch.device(device) if device else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    if dist.get_rank() == src:
        assert blob is not None
        size[0] = int(blob.nbytes)
    dist.broadcast(size, src)
    buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(blob).view(np.uint8)))
    dist.broadcast(buf, src)
    return buf


class LayoutBroadcaster:
    """Layout updates over NCCL inside the library (jg_engine_comm_init / jg_cde_broadcast): rank `root` flattens its host
    mirror and every other rank binds the blob straight from the NCCL receive buffer - one collective and one host wait
    per update on the receiving ranks, none on the root. torch.distributed only hands out the 128-byte NCCL id once."""

    def __init__(self, cde, dist, rank: int, world: int, max_blob_bytes: int = 1 << 20):
        import ctypes as C

        import torch

        from . import JG_OK, JaguaCudaError

        self.cde, self.rank, self.world = cde, int(rank), int(world)
        lib = cde._lib
        nid = np.zeros(128, np.uint8)
        if rank == 0:
            rc = lib.jg_nccl_unique_id(nid)
            if rc != JG_OK:
                raise JaguaCudaError(rc, "jg_nccl_unique_id: libnccl.so.2 could not be loaded")
        if world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            tid = torch.from_numpy(nid).to(dev)
            dist.broadcast(tid, 0)
            nid = np.ascontiguousarray(tid.cpu().numpy())
        # NCCL writes its version banner to stdout when the first communicator of a process is created (NCCL_DEBUG=VERSION
        # and up): keep stdout clean for callers that print machine-readable lines (bench.py) - the banner goes to stderr
        import os
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            rc = lib.jg_engine_comm_init(cde._eng, nid, self.rank, self.world, int(max_blob_bytes))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        cde._chk(rc)
        self._C = C
        self.info = {"transport": "ncclBroadcast inside libjagua_cuda.so (dlopen libnccl.so.2), fixed-capacity exchange buffer",
                     "capacity_bytes": int(max_blob_bytes), "updates": 0, "last_ms": None}

    def broadcast(self, root: int = 0) -> int:
        """Collective: every rank calls it. Returns the layout epoch now bound on this rank."""
        import time

        C = self._C
        ep = C.c_uint64(0)
        t0 = time.perf_counter()
        host = self.cde._cde if self.rank == root else None
        self.cde._chk(self.cde._lib.jg_cde_broadcast(host, self.cde._eng, int(root), C.byref(ep)))
        self.cde._dirty = False
        self.info["updates"] += 1
        self.info["last_ms"] = (time.perf_counter() - t0) * 1e3
        return int(ep.value)


def bind_to_gpu_numa_node(local_rank: int, lib=None) -> dict:
    """Pins this process to the CPU cores of the NUMA node the GPU hangs off, BEFORE any pinned host buffer is
    allocated (first touch puts the pages on the calling thread's node). With 8 ranks each streaming 53 GB/s of
    queries from host memory the staging buffers must sit on the GPU's own socket: SURVEY section 8(e) "check NUMA
    placement of pinned buffers". Reads the PCI address from the CUDA runtime and the topology from sysfs; does
    nothing (and says why) when either is unavailable."""
    import ctypes as C
    import os

    info = {"numa_node": None, "cpus": None, "why": None}
    try:
        rt = None
        for name in ("libcudart.so", "libcudart.so.12"):
            try:
                rt = C.CDLL(name)
                break
            except OSError:
                continue
        if rt is None:
            import glob
            import torch
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
            rt = C.CDLL(cands[0]) if cands else None
        if rt is None:
            info["why"] = "libcudart not found"
            return info
        buf = C.create_string_buffer(32)
        if rt.cudaDeviceGetPCIBusId(buf, 32, int(local_rank)) != 0:
            info["why"] = "cudaDeviceGetPCIBusId failed"
            return info
        bus = buf.value.decode().lower()
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            info["why"] = "sysfs reports no NUMA node for the GPU"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            info["why"] = "none of the node's cores is in this process's affinity mask"
            return info
        os.sched_setaffinity(0, allowed)
        info.update(numa_node=node, cpus=len(allowed))
    except (OSError, ValueError, AttributeError) as e:
        info["why"] = f"{type(e).__name__}: {e}"
    return info
