"""In-process multi-device pool (jg_pool_* in include/jagua_cuda.h): the layout replicated on every GPU of this
process, ONE host batch cut into contiguous shards, every device writing its shard's results into its own slice of the
caller's result array (BASELINE north_star "results are gathered on the host").

    host = jg.build_engine(wl, device=None)          # host mirror of the CDEngine (quadtree maintenance, flatten)
    pool = jg.EnginePool(host, devices=[0, 1, 2, 3])
    pool.upload_items(wl.items); pool.sync_layout()
    verdicts = pool.detect_multi(item_ids, seg_offsets, transforms)   # one array, all devices
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np


class EnginePool:
    def __init__(self, host_cde, devices: Sequence[int]):
        from . import JG_OK, JaguaCudaError, load_library

        self._lib = load_library()
        self._host = host_cde
        self._pool = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self._lib.jg_pool_create(devs, len(devices), C.byref(host_cde.config), C.byref(self._pool))
        if rc != JG_OK:
            raise JaguaCudaError(rc, "jg_pool_create failed: no usable CUDA device (this engine has no CPU fallback)")
        self.devices = [int(d) for d in devices]
        self._epoch_uploaded = None

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_pool", None) and self._pool.value:
            self._lib.jg_pool_destroy(self._pool)
            self._pool = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int) -> int:
        from . import JaguaCudaError

        if rc < 0:
            raise JaguaCudaError(rc, self._lib.jg_pool_last_error(self._pool).decode())
        return rc

    def __len__(self):
        return self._lib.jg_pool_size(self._pool)

    def engine_stats(self, k: int) -> dict:
        a = (C.c_uint64 * 8)()
        self._lib.jg_engine_stats(self._lib.jg_pool_engine(self._pool, int(k)), a)
        return dict(launches=int(a[0]), queries=int(a[1]), h2d_bytes=int(a[2]), d2h_bytes=int(a[3]), layout_uploads=int(a[4]))

    def set_batching(self, chunk_queries: int, n_streams: int):
        for k in range(len(self)):
            rc = self._lib.jg_engine_set_batching(self._lib.jg_pool_engine(self._pool, k), int(chunk_queries), int(n_streams))
            if rc < 0:
                self._chk(rc)

    def set_min_shard(self, queries_per_device: int):
        self._chk(self._lib.jg_pool_set_min_shard(self._pool, int(queries_per_device)))

    # -- replicated state -----------------------------------------------------------------------
    def upload_item(self, item_id: int, shape):
        from . import _f32

        v = _f32(shape.verts).reshape(-1, 2)
        poles = _f32(shape.poles).reshape(-1, 3)
        piers = _f32(shape.piers if shape.piers is not None else np.zeros((0, 4))).reshape(-1, 4)
        self._chk(self._lib.jg_pool_item_upload(self._pool, int(item_id), v, len(v), _f32(shape.poi)[:2].copy(),
                                                poles if len(poles) else np.zeros((1, 3), np.float32), len(poles),
                                                piers if len(piers) else np.zeros((1, 4), np.float32), len(piers)))

    def upload_items(self, items):
        from . import ItemShape

        for i, it in enumerate(items):
            self.upload_item(i, ItemShape(it.verts, it.poi, it.poles, it.piers))

    def set_item_rotations(self, item_id: int, rotations):
        from . import compose

        rots = np.asarray(rotations, np.float32).reshape(-1)
        cs = np.ascontiguousarray(compose(np.stack([rots, np.zeros_like(rots), np.zeros_like(rots)], 1))[:, :2])
        self._chk(self._lib.jg_pool_item_set_rotations(self._pool, int(item_id), cs.ctypes.data, len(rots)))

    def upload_filter(self, filter_id: int, bits: np.ndarray):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        self._chk(self._lib.jg_pool_filter_upload(self._pool, int(filter_id), bits, len(bits)))

    def sync_layout(self, force: bool = False) -> int:
        """Flattens the host mirror once and replicates the blob on every device if it changed since the last upload."""
        ep = self._host.epoch()
        if force or ep != self._epoch_uploaded:
            out = C.c_uint64(0)
            self._chk(self._lib.jg_pool_cde_sync(self._host._cde, self._pool, C.byref(out)))
            self._epoch_uploaded = ep
        return ep

    def upload_layout_blob(self, blob: np.ndarray, epoch: int = 0):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._chk(self._lib.jg_pool_layout_upload(self._pool, blob.ctypes.data, blob.nbytes, int(epoch)))
        self._epoch_uploaded = self._host.epoch()

    # -- queries --------------------------------------------------------------------------------
    @staticmethod
    def _segments(item_ids, seg_offsets):
        ids = np.ascontiguousarray(item_ids, dtype=np.int32)
        offs = np.ascontiguousarray(seg_offsets, dtype=np.int64)
        assert len(offs) == len(ids) + 1 and offs[0] == 0
        return ids, offs

    def detect_multi(self, item_ids, seg_offsets, transforms: np.ndarray, filter_id: int = -1, poly_only: bool = False,
                     out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        assert transforms.dtype == np.float32 and transforms.flags.c_contiguous and transforms.size // 4 == int(offs[-1])
        if out is None:
            out = np.empty(int(offs[-1]), np.uint8)
        fn = self._lib.jg_pool_detect_multi if wait else self._lib.jg_pool_detect_multi_async
        self._chk(fn(self._pool, len(ids), ids.ctypes.data, offs.ctypes.data, transforms.ctypes.data, int(filter_id), 1 if poly_only else 0,
                     out.ctypes.data))
        return out

    def detect_multi_indexed(self, item_ids, seg_offsets, rot_idx: np.ndarray, txy: np.ndarray, filter_id: int = -1,
                             poly_only: bool = False, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        assert rot_idx.dtype == np.uint8 and rot_idx.flags.c_contiguous and txy.dtype == np.float32 and txy.flags.c_contiguous
        assert rot_idx.size == int(offs[-1]) and txy.size == 2 * int(offs[-1])
        if out is None:
            out = np.empty(int(offs[-1]), np.uint8)
        fn = self._lib.jg_pool_detect_multi_indexed if wait else self._lib.jg_pool_detect_multi_indexed_async
        self._chk(fn(self._pool, len(ids), ids.ctypes.data, offs.ctypes.data, rot_idx.ctypes.data, txy.ctypes.data, int(filter_id),
                     1 if poly_only else 0, out.ctypes.data))
        return out

    def detect_multi_lattice(self, item_ids, seg_offsets, lattice, rot_idx: np.ndarray, ij: np.ndarray, filter_id: int = -1,
                             poly_only: bool = False, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        assert rot_idx.dtype == np.uint8 and rot_idx.flags.c_contiguous and ij.dtype == np.uint16 and ij.flags.c_contiguous
        assert rot_idx.size == int(offs[-1]) and ij.size == 2 * int(offs[-1])
        if out is None:
            out = np.empty(int(offs[-1]), np.uint8)
        fn = self._lib.jg_pool_detect_multi_lattice if wait else self._lib.jg_pool_detect_multi_lattice_async
        self._chk(fn(self._pool, len(ids), ids.ctypes.data, offs.ctypes.data, C.addressof(lattice), rot_idx.ctypes.data, ij.ctypes.data,
                     int(filter_id), 1 if poly_only else 0, out.ctypes.data))
        return out

    def collect_multi(self, item_ids, seg_offsets, transforms: np.ndarray, words: Optional[int] = None, poly_only: bool = False,
                      seed: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        self.sync_layout()
        ids, offs = self._segments(item_ids, seg_offsets)
        n = int(offs[-1])
        assert transforms.dtype == np.float32 and transforms.flags.c_contiguous and transforms.size // 4 == n
        words = int(words or self._host.collector_words())
        if out is None:
            out = np.empty((n, words), np.uint32)
        if seed is not None:
            assert seed.dtype == np.uint32 and seed.flags.c_contiguous and seed.shape == (n, words)
        fn = self._lib.jg_pool_collect_multi if wait else self._lib.jg_pool_collect_multi_async
        self._chk(fn(self._pool, len(ids), ids.ctypes.data, offs.ctypes.data, transforms.ctypes.data, 1 if poly_only else 0,
                     seed.ctypes.data if seed is not None else None, out.ctypes.data, words))
        return out

    def sync(self):
        self._chk(self._lib.jg_pool_sync(self._pool))
