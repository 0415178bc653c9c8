//! FFI of `libjagua_cuda` (include/jagua_cuda.h) + the safe wrapper a jagua-rs host uses.
//!
//! NOT COMPILED IN THE BUILD IMAGE OF THIS REPOSITORY (no Rust toolchain there): written against jagua-rs v0.6.4's public
//! API as its own benches use it (`lbf/benches/ci_bench.rs`, `lbf/src/opt/search.rs`). What was built and tested is
//! everything behind the C ABI; the ctypes mirror in `jagua-rs_b200/__init__.py` makes the same calls in the same order.
//!
//! The host keeps its own `CDEngine` (hazard `SlotMap`, quadtree, snapshots). After every `register_hazard` /
//! `deregister_hazard_*` / `restore` / `swap_container` it calls [`GpuCde::sync_layout`], which flattens the `QTNode` tree
//! ([`flatten::flatten_cde`]) and uploads it; queries are then whole batches:
//!
//! ```ignore
//! let gpu = GpuCde::new(0, cde.config)?;                       // one per CUDA device
//! gpu.upload_item(item)?;                                      // Item::shape_cd + its SPSurrogate
//! gpu.sync_layout(layout.cde())?;                              // after every layout change
//! let verdicts = gpu.detect_batch(item, &transformations, None)?;      // detect_surrogate_collision + detect_poly_collision
//! let sets = gpu.collect_batch(item, &transformations, layout.cde())?; // Vec<BasicHazardCollector>
//! ```
#![allow(non_camel_case_types)]

pub mod flatten;

use jagua_rs::collision_detection::hazards::collector::{BasicHazardCollector, HazardCollector};
use jagua_rs::collision_detection::hazards::HazKey;
use jagua_rs::collision_detection::{CDEConfig, CDEngine};
use jagua_rs::entities::Item;
use jagua_rs::geometry::Transformation;
use slotmap::Key;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct JgEngine {
    _p: [u8; 0],
}
#[repr(C)]
pub struct JgPool {
    _p: [u8; 0],
}
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct JgConfig {
    pub quadtree_depth: i32,
    pub cd_threshold: i32,
    pub n_ff_poles: i32,
    pub n_ff_piers: i32,
}
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct JgLattice {
    pub x0: f32,
    pub y0: f32,
    pub dx: f32,
    pub dy: f32,
}

pub const JG_OK: c_int = 0;
pub const JG_VERDICT_SURROGATE: u8 = 1;
pub const JG_VERDICT_POLY: u8 = 2;
/// the CPU path would hit `unreachable!` at cd_engine.rs:179 for this query (an Entire hazard in the virtual root hidden by the filter)
pub const JG_VERDICT_REF_PANIC: u8 = 0x80;
pub const JG_MODE_DEFAULT: c_int = 0;
pub const JG_MODE_POLY_ONLY: c_int = 1;

extern "C" {
    pub fn jg_engine_create(device: c_int, cfg: *const JgConfig, out: *mut *mut JgEngine) -> c_int;
    pub fn jg_engine_destroy(e: *mut JgEngine);
    pub fn jg_last_error(e: *const JgEngine) -> *const c_char;
    pub fn jg_layout_upload(e: *mut JgEngine, blob: *const u8, bytes: usize, epoch: u64) -> c_int;
    pub fn jg_layout_upload_device(e: *mut JgEngine, d_blob: *const u8, bytes: usize, epoch: u64) -> c_int;
    pub fn jg_nccl_unique_id(id: *mut u8) -> c_int;
    pub fn jg_engine_comm_init(e: *mut JgEngine, id: *const u8, rank: c_int, world: c_int, max_blob_bytes: usize) -> c_int;
    pub fn jg_layout_broadcast(e: *mut JgEngine, root: c_int, blob: *const u8, bytes: usize, epoch: u64) -> c_int;
    pub fn jg_item_upload(e: *mut JgEngine, item_id: c_int, verts_xy: *const f32, n_verts: c_int, poi_xy: *const f32,
                          poles_xyr: *const f32, n_poles: c_int, piers: *const f32, n_piers: c_int) -> c_int;
    pub fn jg_item_set_rotations(e: *mut JgEngine, item_id: c_int, rot_cos_sin: *const f32, n_rot: c_int) -> c_int;
    pub fn jg_filter_upload(e: *mut JgEngine, filter_id: c_int, bits: *const u32, words: c_int) -> c_int;
    pub fn jg_filter_upload_keys(e: *mut JgEngine, filter_id: c_int, hkeys: *const u64, n_keys: c_int) -> c_int;
    pub fn jg_detect_batch(e: *mut JgEngine, item_id: c_int, transforms: *const f32, n: i64, filter_id: c_int, mode: c_int,
                           out: *mut u8) -> c_int;
    pub fn jg_detect_multi(e: *mut JgEngine, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64, transforms: *const f32,
                           filter_id: c_int, mode: c_int, out: *mut u8) -> c_int;
    pub fn jg_detect_multi_async(e: *mut JgEngine, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64,
                                 transforms: *const f32, filter_id: c_int, mode: c_int, out: *mut u8) -> c_int;
    pub fn jg_detect_multi_indexed_async(e: *mut JgEngine, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64,
                                         rot_idx: *const u8, txy: *const f32, filter_id: c_int, mode: c_int, out: *mut u8) -> c_int;
    pub fn jg_detect_multi_lattice_async(e: *mut JgEngine, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64,
                                         lattice: *const JgLattice, rot_idx: *const u8, ij: *const u16, filter_id: c_int,
                                         mode: c_int, out: *mut u8) -> c_int;
    pub fn jg_collect_multi(e: *mut JgEngine, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64, transforms: *const f32,
                            mode: c_int, seed_bits: *const u32, out_bits: *mut u32, words: c_int) -> c_int;
    pub fn jg_sync(e: *mut JgEngine) -> c_int;
    pub fn jg_host_alloc(bytes: usize, out: *mut *mut u8) -> c_int;
    pub fn jg_host_free(p: *mut u8) -> c_int;
    // in-process multi-device pool: one host batch sharded over the GPUs of the box, results gathered into one array
    pub fn jg_pool_create(devices: *const c_int, n_devices: c_int, cfg: *const JgConfig, out: *mut *mut JgPool) -> c_int;
    pub fn jg_pool_destroy(p: *mut JgPool);
    pub fn jg_pool_layout_upload(p: *mut JgPool, blob: *const u8, bytes: usize, epoch: u64) -> c_int;
    pub fn jg_pool_item_upload(p: *mut JgPool, item_id: c_int, verts_xy: *const f32, n_verts: c_int, poi_xy: *const f32,
                               poles_xyr: *const f32, n_poles: c_int, piers: *const f32, n_piers: c_int) -> c_int;
    pub fn jg_pool_detect_multi(p: *mut JgPool, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64, transforms: *const f32,
                                filter_id: c_int, mode: c_int, out: *mut u8) -> c_int;
    pub fn jg_pool_collect_multi(p: *mut JgPool, n_seg: c_int, item_ids: *const c_int, seg_offsets: *const i64, transforms: *const f32,
                                 mode: c_int, seed_bits: *const u32, out_bits: *mut u32, words: c_int) -> c_int;
    pub fn jg_pool_sync(p: *mut JgPool) -> c_int;
    pub fn jg_pool_last_error(p: *const JgPool) -> *const c_char;
}

#[derive(Debug)]
pub struct GpuError(pub c_int, pub String);
pub type Result<T> = std::result::Result<T, GpuError>;

/// Device half of a `CDEngine`: owns the HBM copies of one flattened layout, the item shapes and the filters.
pub struct GpuCde {
    raw: *mut JgEngine,
    epoch: u64,
    words: usize,
}
// calls on one handle are serialised by &mut / the caller; different handles are independent
unsafe impl Send for GpuCde {}

/// id handed to `detect_batch` for a filter uploaded with [`GpuCde::upload_filter_keys`]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct FilterId(pub i32);

impl GpuCde {
    pub fn new(device: i32, cfg: CDEConfig) -> Result<Self> {
        let c = JgConfig {
            quadtree_depth: cfg.quadtree_depth as i32,
            cd_threshold: cfg.cd_threshold as i32,
            n_ff_poles: cfg.item_surrogate_config.n_ff_poles as i32,
            n_ff_piers: cfg.item_surrogate_config.n_ff_piers as i32,
        };
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { jg_engine_create(device, &c, &mut raw) };
        if rc != JG_OK {
            return Err(GpuError(rc, "jg_engine_create: no usable CUDA device (there is no CPU fallback)".into()));
        }
        Ok(GpuCde { raw, epoch: 0, words: 1 })
    }

    fn check(&self, rc: c_int) -> Result<()> {
        if rc == JG_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(jg_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(GpuError(rc, msg))
    }

    /// `Item::shape_cd` (vertices, poi) + its `SPSurrogate` (poles as (x, y, r), piers as (sx, sy, ex, ey)); the engine tests
    /// the first `n_ff_poles` poles and all piers (sp_surrogate.rs:55-61).
    pub fn upload_item(&mut self, item: &Item) -> Result<()> {
        let shape = item.shape_cd.as_ref();
        let verts: Vec<f32> = shape.vertices.iter().flat_map(|p| [p.0, p.1]).collect();
        let poi = [shape.poi.center.0, shape.poi.center.1];
        let surr = shape.surrogate();
        let poles: Vec<f32> = surr.poles.iter().flat_map(|c| [c.center.0, c.center.1, c.radius]).collect();
        let piers: Vec<f32> = surr.ff_piers().iter().flat_map(|e| [e.start.0, e.start.1, e.end.0, e.end.1]).collect();
        let rc = unsafe {
            jg_item_upload(self.raw, item.id as c_int, verts.as_ptr(), (verts.len() / 2) as c_int, poi.as_ptr(), poles.as_ptr(),
                           (poles.len() / 3) as c_int, piers.as_ptr(), (piers.len() / 4) as c_int)
        };
        self.check(rc)
    }

    /// Re-flatten `cde` and upload it. Stream-ordered: batches in flight finish on the previous layout.
    pub fn sync_layout(&mut self, cde: &CDEngine) -> Result<()> {
        let blob = flatten::flatten_cde(cde);
        self.epoch += 1;
        self.words = flatten::collector_words(cde);
        let rc = unsafe { jg_layout_upload(self.raw, blob.as_ptr(), blob.len(), self.epoch) };
        self.check(rc)
    }

    /// `HazKeyFilter` (hazards/filter.rs:12-58) by versioned keys: stays valid across layout changes (a stale key matches nothing).
    pub fn upload_filter_keys(&mut self, id: FilterId, keys: impl Iterator<Item = HazKey>) -> Result<()> {
        let k: Vec<u64> = keys.map(|k| k.data().as_ffi()).collect();
        let rc = unsafe { jg_filter_upload_keys(self.raw, id.0, k.as_ptr(), k.len() as c_int) };
        self.check(rc)
    }

    /// Batched `detect_surrogate_collision` + `detect_poly_collision` (cd_engine.rs:160-223) for one item under many
    /// transformations. Verdict byte: bit0 surrogate, bit1 polygon, bit7 = the CPU path would panic (N3).
    pub fn detect_batch(&mut self, item: &Item, transfs: &[Transformation], filter: Option<FilterId>) -> Result<Vec<u8>> {
        let t4 = compose4(transfs);
        let mut out = vec![0u8; transfs.len()];
        let rc = unsafe {
            jg_detect_batch(self.raw, item.id as c_int, t4.as_ptr() as *const f32, t4.len() as i64, filter.map_or(-1, |f| f.0),
                            JG_MODE_DEFAULT, out.as_mut_ptr())
        };
        self.check(rc)?;
        Ok(out)
    }

    /// Batched `collect_surrogate_collisions` + `collect_poly_collisions` (cd_engine.rs:261-310) into one collector per
    /// transformation; bits (hazard slot indices) are mapped back to `(HazKey, HazardEntity)` through `cde.hazards_map`.
    pub fn collect_batch(&mut self, item: &Item, transfs: &[Transformation], cde: &CDEngine) -> Result<Vec<BasicHazardCollector>> {
        let t4 = compose4(transfs);
        let words = self.words;
        let mut bits = vec![0u32; transfs.len() * words];
        let offs = [0i64, transfs.len() as i64];
        let id = item.id as c_int;
        let rc = unsafe {
            jg_collect_multi(self.raw, 1, &id, offs.as_ptr(), t4.as_ptr() as *const f32, JG_MODE_DEFAULT, std::ptr::null(),
                             bits.as_mut_ptr(), words as c_int)
        };
        self.check(rc)?;
        let by_slot = flatten::keys_by_slot(cde);
        Ok(bits
            .chunks(words)
            .map(|set| {
                let mut c = BasicHazardCollector::with_capacity(8);
                for (w, &word) in set.iter().enumerate() {
                    let mut m = word;
                    while m != 0 {
                        let slot = w * 32 + m.trailing_zeros() as usize;
                        m &= m - 1;
                        if let Some(hkey) = by_slot.get(slot).copied().flatten() {
                            c.insert(hkey, cde.hazards_map[hkey].entity);
                        }
                    }
                }
                c
            })
            .collect())
    }

    pub fn sync(&mut self) -> Result<()> {
        let rc = unsafe { jg_sync(self.raw) };
        self.check(rc)
    }
}

impl Drop for GpuCde {
    fn drop(&mut self) {
        unsafe { jg_engine_destroy(self.raw) }
    }
}

/// `Transformation` = [[cos, -sin, tx], [sin, cos, ty], [0, 0, 1]] (geometry/transformation.rs:11-13,91-102): the four meaningful
/// entries, composed on the host by `f32::sin_cos` exactly as the CPU path composes them. The device never evaluates sin / cos.
fn compose4(transfs: &[Transformation]) -> Vec<[f32; 4]> {
    transfs
        .iter()
        .map(|t| {
            let m = t.matrix();
            [m[0][0].into_inner(), m[1][0].into_inner(), m[0][2].into_inner(), m[1][2].into_inner()]
        })
        .collect()
}
