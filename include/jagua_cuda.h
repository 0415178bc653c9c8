/* jagua_cuda.h — C ABI of the B200-native batched collision-query engine for jagua-rs's CDE.
 *
 * The reference (jagua-rs v0.6.4, pure Rust) has no FFI seam; the seam is the Rust API of
 * `CDEngine` (jagua-rs/src/collision_detection/cd_engine.rs:33-342) plus the `HazardFilter`
 * (hazards/filter.rs:7-9) and `HazardCollector` (hazards/collector.rs:7-42) traits. These entry
 * points are what a `jagua-cuda-sys` crate would bind: plain pointers and
 * sizes only. All functions return JG_OK (0) or a negative JgStatus; none of them aborts.
 *
 * Two layers:
 *   jg_engine_*  device engine: owns HBM copies of one flattened layout (quadtree + hazards), the
 *                item shapes and the filters, and resolves batches of queries with sm_100a kernels.
 *   jg_cde_*     host-side mirror of CDEngine's update path (register / deregister / save /
 *                restore and the quadtree maintenance behind them) that produces the flattened
 *                layout blob. A Rust host keeps its own CDEngine and only needs jg_engine_*.
 */
#ifndef JAGUA_CUDA_H
#define JAGUA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct JgEngine JgEngine;
typedef struct JgCde JgCde;

typedef enum JgStatus {
    JG_OK = 0,
    JG_ERR_CUDA = -1,         /* a CUDA runtime call failed; see jg_last_error */
    JG_ERR_INVALID = -2,      /* bad argument / unknown id / malformed blob */
    JG_ERR_NO_LAYOUT = -3,    /* query before any layout upload */
    JG_ERR_CAPACITY = -4,     /* e.g. collector words too small for the layout's hazard slots */
    JG_ERR_REFERENCE_PANIC = -5 /* host mirror reached a state where the reference would panic */
} JgStatus;

/* CDEConfig (cd_engine.rs:346-353) + SPSurrogateConfig's fail-fast counts (sp_surrogate.rs:120-131) */
typedef struct JgConfig {
    int32_t quadtree_depth; /* CDEConfig::quadtree_depth */
    int32_t cd_threshold;   /* CDEConfig::cd_threshold (collect only, qt_node.rs:157) */
    int32_t n_ff_poles;     /* SPSurrogateConfig::n_ff_poles */
    int32_t n_ff_piers;     /* SPSurrogateConfig::n_ff_piers (ff_piers() = all piers uploaded) */
} JgConfig;

/* HazardEntity discriminants (hazards/hazard.rs:36-53) */
enum { JG_HAZ_PLACED_ITEM = 0, JG_HAZ_EXTERIOR = 1, JG_HAZ_HOLE = 2, JG_HAZ_QZONE = 3 };

/* detect verdict byte (one per query):
 *   bit0  detect_surrogate_collision == true                (cd_engine.rs:204-223)
 *   bit1  detect_poly_collision == true, evaluated only when bit0 == 0 (cd_engine.rs:160-197; the
 *         two-stage order of lbf/src/opt/search.rs:44-57 and lbf/benches/ci_bench.rs:107-121)
 *   bit7  the reference would hit `unreachable!` at cd_engine.rs:179 for this query (an Entire
 *         hazard in the virtual root that the filter hides); bit1 is 0 then.                    */
enum { JG_VERDICT_SURROGATE = 1, JG_VERDICT_POLY = 2, JG_VERDICT_REF_PANIC = 0x80 };

/* query modes */
enum {
    JG_MODE_DEFAULT = 0,
    JG_MODE_POLY_ONLY = 1 /* skip the surrogate stage (lbf/benches/quadtree_bench.rs:73-124) */
};

/* ------------------------------------------------------------------------------------------ */
/* Limits (results never depend on which kernels serve a batch; speed does)                    */
/* ------------------------------------------------------------------------------------------ */
/* quadtree depth            <= 15 (JgConfig::quadtree_depth; the walk state keeps 4 pending bits per level)
 * quadtree nodes            <= 2^20 (JG_ERR_CAPACITY beyond)
 * item vertices             <= 4096
 * queries per launch piece  <= 2^26 (larger batches are cut into pieces)
 * collector words           <= 1024 (32768 hazard slots)
 * GRID kernels (k_s1_grid / k_s2_grid / k_collect_grid: the fast path, 2-4x the walk kernels) serve a layout when
 *   - the quadtree is at most 5 deep (the reference's default, lbf/src/config.rs:33-41) and has <= 0xFFFE nodes,
 *   - every node box equals the nested-midpoint grid bit for bit (true for every tree built by CDEngine::new),
 *   - no leaf holds more than 255 Partial hazards,
 *   and, for hazard-set collection, when in addition
 *   - the collector has <= 8 words (256 hazard slots),
 *   - items carry <= 4 fail-fast poles, <= 4 piers and <= 96 vertices.
 * Everything else runs on the quadtree-walk kernels (any depth <= 15, any item). jg_engine_detect_path reports which.
 * Stage 2 stages the layout's entries + edges in shared memory by TMA when they fit 220 KB (swim: 100 KB), else reads
 * them through L1. */

/* ------------------------------------------------------------------------------------------ */
/* Device engine                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* CDEngine::new's device half: binds to CUDA device `device` and creates the engine's streams. */
int jg_engine_create(int device, const JgConfig* cfg, JgEngine** out);
void jg_engine_destroy(JgEngine* e);
const char* jg_last_error(const JgEngine* e);

/* Replaces the engine's layout (quadtree nodes, hazard entries, edges, hazard shapes) with the
 * flattened blob produced by jg_cde_flatten (format: JgLayoutHeader in layout_soa.h). Mirrors the
 * effect of CDEngine::{register_hazard, deregister_hazard_*, restore} (cd_engine.rs:59-149) on the
 * data the queries read. `epoch` is returned by jg_engine_layout_epoch; queries issued after this
 * call see the new layout, queries in flight finish on the old one. Host pointer. */
int jg_layout_upload(JgEngine* e, const void* blob, size_t bytes, uint64_t epoch);
/* Incremental update (BASELINE north_star item 1 "updated incrementally on register/deregister"):
 * `image` is the CURRENT state of a position-stable blob (jg_cde_image_sync: node records at fixed indices, entry / edge /
 * vertex blocks in section heaps) and ranges_off_len holds n_ranges (byte offset, byte length) pairs that changed since
 * sequence number seq - 1 of the same lineage. The engine copies only what its target buffer is missing - as ONE patch
 * buffer plus a scatter kernel - or the whole image when it does not hold this lineage (first use, a missed sync, a rebuilt
 * image). Stream-ordered like jg_layout_upload. The image comes from this library's own host mirror and is not re-validated. */
int jg_layout_patch(JgEngine* e, const void* image, size_t bytes, uint64_t lineage, uint64_t seq, const uint32_t* ranges_off_len, int n_ranges,
                    uint64_t epoch);
/* Same, blob already in this device's memory (e.g. the receive buffer of an NCCL broadcast). */
int jg_layout_upload_device(JgEngine* e, const void* d_blob, size_t bytes, uint64_t epoch);
uint64_t jg_engine_layout_epoch(const JgEngine* e);

/* Layout updates over NCCL (BASELINE north_star: "the layout is small, so it is replicated on every GPU ... NCCL over
 * NVLink is used only to broadcast layout updates"): one engine per rank (process or thread), jg_nccl_unique_id on one
 * rank, the 128 id bytes handed to every rank by whatever the host has (torch.distributed, MPI, a pipe),
 * jg_engine_comm_init on every rank. max_blob_bytes is the capacity of the exchange buffer (0: 1 MiB); every update
 * broadcasts exactly that many bytes, because receivers learn a blob's size only from its header. NCCL is loaded at run
 * time (dlopen "libnccl.so.2": the copy a process already holds - torch's - is shared); engines that never call
 * jg_engine_comm_init do not need it.
 * jg_layout_broadcast is a collective: every rank calls it, in the same order. The root passes the flattened layout (host
 * pointer) and gets it by the ordinary stream-ordered upload; the other ranks pass blob = NULL, wait ONCE for the
 * arrival (the header and the device-side structural check come back together) and bind the blob from device memory.
 * Replaces what Layout::place_item / remove_item / restore do to the CDE (entities/layout.rs:52-100) on the other GPUs.
 * Queries in flight on any rank finish on the old layout. jg_cde_broadcast = flatten on the root + jg_layout_broadcast
 * (c may be NULL on the other ranks). */
int jg_nccl_unique_id(uint8_t id[128]);
int jg_engine_comm_init(JgEngine* e, const uint8_t id[128], int rank, int world, size_t max_blob_bytes);
int jg_engine_comm_destroy(JgEngine* e);
int jg_layout_broadcast(JgEngine* e, int root, const void* blob, size_t bytes, uint64_t epoch);
int jg_cde_broadcast(JgCde* c, JgEngine* e, int root, uint64_t* epoch);

/* An item's collision shape: SPolygon.vertices (x,y pairs, CCW, already centred), SPolygon.poi
 * centre (simple_polygon.rs:24-37), SPSurrogate.poles (x,y,r) and piers (sx,sy,ex,ey)
 * (sp_surrogate.rs:17-30). The first min(n_ff_poles, n_poles) poles and all piers are tested. */
int jg_item_upload(JgEngine* e, int item_id, const float* verts_xy, int n_verts, const float* poi_xy,
                   const float* poles_xyr, int n_poles, const float* piers, int n_piers);

/* HazardFilter (hazards/filter.rs:7-76) as a bitset over hazard slot indices: bit set = the hazard
 * is irrelevant. NoFilter = filter_id -1 in the query calls. A collector used as a filter (the blanket impl,
 * filter.rs:67-76) is its bitset uploaded here. A slot bitset does not carry HazKey versions: the hazard SlotMap
 * reuses slots (LIFO), so a bitset filter describes ONE layout - after any register / deregister upload it again,
 * or use jg_filter_upload_keys, which the engine re-derives against every layout it receives. */
int jg_filter_upload(JgEngine* e, int filter_id, const uint32_t* bits, int words);
/* HazKeyFilter (hazards/filter.rs:12-58) by versioned keys (slot index | version << 32, as jg_cde_register_hazard returns
 * them): a key matches a slot only while that slot still holds that version, exactly like the reference's SecondaryMap
 * lookup - a stale key never hides the hazard that reuses its slot. The bitset is rebuilt from the keys against the slot
 * versions of every layout upload (layout blob section haz_version). n_keys may be 0 (an empty filter). */
int jg_filter_upload_keys(JgEngine* e, int filter_id, const uint64_t* hkeys, int n_keys);

/* detect_surrogate_collision + detect_poly_collision for n transformations of one item.
 * transforms: n x (cos, sin, tx, ty) — the meaningful entries of `Transformation`
 * (transformation.rs:11-13,131-139) composed on the host; out: n verdict bytes. HOST pointers
 * (pinned memory from jg_host_alloc overlaps upload, kernels and download across chunks). */
int jg_detect_batch(JgEngine* e, int item_id, const float* transforms, int64_t n, int filter_id, int mode,
                    uint8_t* out);
/* Same, but returns once every chunk is enqueued; `transforms` and `out` must stay valid (and should
 * be pinned) until jg_sync. Consecutive calls pipeline across the engine's chunk streams. */
int jg_detect_batch_async(JgEngine* e, int item_id, const float* transforms, int64_t n, int filter_id, int mode,
                          uint8_t* out);
/* Multi-item batch: segment k = queries seg_offsets[k] .. seg_offsets[k+1] (seg_offsets[0] == 0) place
 * item item_ids[k]; one launch sequence resolves all segments (the per-item calls are the 1-segment
 * case). This is how a caller that samples several item types at once (sparrow, or lbf over its item
 * list) amortises launches. Host pointers / async / device pointers as above. */
int jg_detect_multi(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                    int filter_id, int mode, uint8_t* out);
int jg_detect_multi_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets,
                          const float* transforms, int filter_id, int mode, uint8_t* out);
/* Compact query form for items with a discrete set of allowed orientations (geo_enums.rs RotationRange::Discrete,
 * sampled by lbf/src/samplers/rotation_distr.rs:42-48): the caller registers, per item, (cos, sin) of every allowed
 * rotation - composed on the host with the platform libm exactly like transformation.rs:131-139 - and a query is
 * (index of the rotation: 1 byte, translation: 2 floats) = 9 B instead of 16 B over PCIe. The kernels rebuild the
 * same (cos, sin, tx, ty) record, so verdicts are identical to jg_detect_multi on the composed transforms.
 * rot_cos_sin: n_rot pairs, 1 <= n_rot <= 256. Indices >= n_rot are clamped to n_rot - 1. */
int jg_item_set_rotations(JgEngine* e, int item_id, const float* rot_cos_sin, int n_rot);
int jg_detect_multi_indexed(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                            const float* txy, int filter_id, int mode, uint8_t* out);
int jg_detect_multi_indexed_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets,
                                  const uint8_t* rot_idx, const float* txy, int filter_id, int mode, uint8_t* out);
int jg_detect_multi_device(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets,
                           const float* d_transforms, int filter_id, int mode, uint8_t* d_out);
/* Same with DEVICE pointers; asynchronous on the engine's compute stream (jg_sync to wait). */
int jg_detect_batch_device(JgEngine* e, int item_id, const float* d_transforms, int64_t n, int filter_id,
                           int mode, uint8_t* d_out);

/* Lattice query form for samplers that draw translations from a regular lattice over their sampling box (a
 * UniformRectSampler, lbf/src/samplers/uniform_rect_sampler.rs:18-45, quantised to 65536 steps per axis): a query is
 * (index of the rotation: 1 byte, lattice coordinates i, j: 2 x u16) = 5 B instead of 16 B over PCIe. Its translation is
 * tx = x0 + (float)i * dx, ty = y0 + (float)j * dy, every product and sum rounded once in f32 (no FMA) - jg_lattice_points
 * is the host twin, so caller and device agree on the transformation bit for bit, and verdicts equal jg_detect_multi on
 * the composed (cos, sin, tx, ty) records. Items need jg_item_set_rotations. */
typedef struct JgLattice {
    float x0, y0, dx, dy;
} JgLattice;
int jg_detect_multi_lattice(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                            const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out);
int jg_detect_multi_lattice_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                                  const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out);
void jg_lattice_points(const JgLattice* lattice, const uint16_t* ij, int64_t n, float* txy);

/* collect_surrogate_collisions + collect_poly_collisions into one collector per query
 * (cd_engine.rs:261-310; lbf/benches/ci_bench.rs:56-70). out_bits: n x words u32, bit = hazard slot
 * index (an order-normalised BasicHazardCollector). words*32 must cover the layout's slot capacity. */
int jg_collect_batch(JgEngine* e, int item_id, const float* transforms, int64_t n, int mode,
                     uint32_t* out_bits, int words);
int jg_collect_batch_async(JgEngine* e, int item_id, const float* transforms, int64_t n, int mode,
                           uint32_t* out_bits, int words);
int jg_collect_batch_device(JgEngine* e, int item_id, const float* d_transforms, int64_t n, int mode,
                            uint32_t* d_out_bits, int words);
/* Multi-item collect batch (segments as in jg_detect_multi): ONE launch resolves all segments. seed_bits (n x words, or
 * NULL for empty collectors) pre-seeds every query's collector: hazards already in a collector are skipped by the
 * traversal (`collector.contains_key`, qt_node.rs:172; cd_engine.rs:279) and stay in the result, which is the reference's
 * behaviour for a collector that is passed in non-empty. seed_bits may alias out_bits (in place). */
int jg_collect_multi(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                     const uint32_t* seed_bits, uint32_t* out_bits, int words);
int jg_collect_multi_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                           const uint32_t* seed_bits, uint32_t* out_bits, int words);
int jg_collect_multi_device(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* d_transforms,
                            int mode, const uint32_t* d_seed_bits, uint32_t* d_out_bits, int words);

/* ------------------------------------------------------------------------------------------ */
/* Batched placement search rounds                                                             */
/* ------------------------------------------------------------------------------------------ */

/* One decision point of lbf::opt::search (lbf/src/opt/search.rs:17-106) as a batch: n samples are DRAWN ON THE DEVICE
 * (mode 0: UniformRectSampler over `box`, rotation uniform over the item's allowed set or fixed to rot_idx >= 0,
 * uniform_rect_sampler.rs:18-45; mode 1: LSSampler, normal(mean, sd) translations around the reference transformation
 * with its rotation rot_idx, ls_sampler.rs:34-120), resolved by the detect kernels (surrogate + polygon stage, under
 * filter_id), and reduced to the free sample of minimal LBFLoss::cost (x_max * 10 + y_max of the placed shape's bbox,
 * lbf/src/opt/loss.rs:12-35). Items need jg_item_set_rotations (a single (cos, sin) = (1, 0) entry if they do not
 * rotate); continuous rotations are not drawn here (their (cos, sin) must come from the host libm). Synchronous; the
 * samples are a pure function of (seed, counter, sample index) - Philox4x32-10, not the reference's SmallRng stream. */
typedef struct JgSearchRound {
    int mode;      /* 0 uniform, 1 normal */
    int rot_idx;   /* fixed rotation index; -1 (mode 0 only): uniform over the allowed rotations */
    float box[4];  /* mode 0: x_min, y_min, x_max, y_max of the sampling box */
    float mean[2]; /* mode 1: reference translation */
    float sd;      /* mode 1: standard deviation of both translation components */
    uint64_t seed, counter;
} JgSearchRound;
typedef struct JgSearchBest {
    int64_t n_free; /* samples without a collision */
    int64_t index;  /* sample index of the best free sample, -1 if none */
    float loss, tx, ty;
    int rot_idx;
} JgSearchBest;
int jg_search_round(JgEngine* e, int item_id, const JgSearchRound* round, int64_t n, int filter_id, JgSearchBest* best);
/* the samples and verdict bytes of the last jg_search_round (tests, debugging): n must match; null pointers are skipped */
int jg_search_samples(JgEngine* e, int64_t n, uint8_t* rot_idx, float* txy, uint8_t* verdicts);

/* Waits for everything enqueued on the engine. Deferred checks surface here: a device-side error of a streamed batch, and a
 * disagreement between the producer's grid premise (JG_LAYOUT_FLAG_GRID_OK in the blob header) and k_grid_build's own verdict
 * for a stream-ordered upload. When jg_sync returns an error, the results of every batch enqueued since the previous successful
 * jg_sync are INVALID and must be discarded (jg_last_error says why); the engine itself stays usable. */
int jg_sync(JgEngine* e);
/* Finer than jg_sync: jg_fence_record marks everything enqueued on this engine so far, jg_fence_wait blocks until that
 * much has completed (later batches keep running). Fences are recycled after 32 newer ones (then a wait degrades to
 * jg_sync). */
int jg_fence_record(JgEngine* e, uint64_t* fence);
int jg_fence_wait(JgEngine* e, uint64_t fence);

/* Counters since engine creation: [0] kernel launches, [1] queries resolved, [2] H2D bytes,
 * [3] D2H bytes, [4] layout uploads, [5] survivors of the surrogate stage and [6] work items that
 * took the register-only fallback walk in the most recent *_device detect call, [7] layout uploads
 * that travelled as patches (jg_layout_patch). */
int jg_engine_stats(const JgEngine* e, uint64_t out[8]);
/* Device time (ms, CUDA events on the compute stream) of the kernels of the most recent
 * *_batch_device call; valid after jg_sync. */
float jg_engine_last_kernel_ms(JgEngine* e);
/* CUDA-event stopwatch on the compute stream (the stream the *_device calls launch on):
 * start records an event, stop records another, waits for it and returns the elapsed ms. */
int jg_timer_start(JgEngine* e);
/* The same with a gate: the compute stream is held until jg_timer_stop has been called, so that everything enqueued in between
 * runs back to back on the device and the host's enqueue time is outside the measured interval. */
int jg_timer_start_gated(JgEngine* e);
int jg_timer_stop(JgEngine* e, float* ms);
/* How the host-pointer batch calls cut a batch into pieces (piece k runs on stream k mod n_streams: H2D, kernels, D2H).
 * chunk_queries == 0 (the default, with 4 streams): chosen per batch - one piece up to 512 Ki queries; engine idle (one batch
 * at a time): pieces of 256 Ki, so that copies and kernels overlap; batches already in flight: whole pieces of up to 1 Mi
 * (their copies hide behind the kernels of the previous batches, one launch pair per piece).
 * chunk_queries > 0 pins fixed pieces of that size. */
int jg_engine_set_batching(JgEngine* e, int64_t chunk_queries, int n_streams);
/* Lanes per query (0 = choose from the item's vertex count, else 4, 8, 16 or 32) of the detect
 * stage-2 and collect kernels. A tuning knob; results do not depend on it. */
int jg_engine_set_groups(JgEngine* e, int detect_group, int collect_group);
/* Launches of at least min_queries queries run detect_containment_collision (cd_engine.rs:231-255) in a kernel of its
 * own after the edge tests (default 1.5 * 2^20; 0 = always, a huge value = never). Results do not depend on it. */
int jg_engine_set_split_min(JgEngine* e, int64_t min_queries);
/* Stage 2 of the grid detect path as a programmatic dependent launch behind stage 1: its CTAs take the SM slots stage 1
 * frees while it drains, stage their tables and wait for stage 1's completion on the device (griddepcontrol.wait).
 * Process-wide, on by default (environment variable JG_PDL=0 or this call restore ordinary stream order; measured: 4 us
 * per launch pair, 59 -> 55 us at 64 Ki queries, 226 -> 222 us at 1 Mi). Results do not depend on it. */
int jg_set_dependent_launch(int on);
int jg_dependent_launch(void);
/* Which detect kernels serve the next layout uploads: JG_PATH_AUTO (grid kernels when the quadtree is at
 * most 5 deep, else the quadtree-walk kernels) or JG_PATH_WALK (always
 * the walk kernels). Takes effect immediately for the current layout. Results do not depend on it. */
#define JG_PATH_AUTO 0
#define JG_PATH_WALK 1
/* JG_PATH_FUSED (experimental): where the grid kernels apply, ONE persistent kernel per batch runs both stages
 * (every warp resolves its own survivors) and, for host batches of >= 128 Ki queries, follows the H2D copy front
 * through an arrival counter instead of waiting for whole chunks. Same results; measured slower than the
 * two-kernel sequence on B200 (the merged instruction footprint), so it is not the default. */
#define JG_PATH_FUSED 2
int jg_engine_set_detect_path(JgEngine* e, int path);
/* 0: quadtree-walk kernels, 1: grid kernels, 2: fused grid kernel (JG_PATH_FUSED). */
int jg_engine_detect_path(const JgEngine* e);

/* Memory helpers (pinned host / device) so callers can keep the C ABI free of CUDA headers. */
int jg_host_alloc(size_t bytes, void** out);
/* write-combined pinned memory (cudaHostAllocWriteCombined): for buffers the host only writes and the device only reads
 * (query records); not snooped on the way over PCIe. Reading it from the host is slow. */
int jg_host_alloc_wc(size_t bytes, void** out);
int jg_host_free(void* p);
int jg_device_alloc(JgEngine* e, size_t bytes, void** out);
int jg_device_free(JgEngine* e, void* p);
int jg_memcpy_h2d(JgEngine* e, void* d_dst, const void* src, size_t bytes);
int jg_memcpy_d2h(JgEngine* e, void* dst, const void* d_src, size_t bytes);

/* DTransformation (rotation, tx, ty) -> (cos, sin, tx, ty) with the host libm's sinf/cosf, which is
 * what f32::sin_cos lowers to (transformation.rs:131-139). The device never evaluates sin/cos. */
void jg_compose_batch(const float* rot_tx_ty, int64_t n, float* out4, int n_threads);

/* Non-FMA FP32 pipe peak of this device (independent FADD/FMUL chains), in Gflop/s: the roofline
 * denominator for kernels compiled with -fmad=false. */
int jg_fp32_nofma_peak(JgEngine* e, double* gflops, float* ms);
/* Writes a buffer larger than L2 (bench hygiene between timed iterations). */
int jg_flush_l2(JgEngine* e);

/* ------------------------------------------------------------------------------------------ */
/* Item preparation (host, once per item type)                                                 */
/* ------------------------------------------------------------------------------------------ */

/* SPSurrogateConfig (geometry/fail_fast/sp_surrogate.rs:120-131); LBF defaults (lbf/src/config.rs:33-41):
 * n_pole_limits {100, 20, 10}, coverage_limits {0.0, 0.75, 0.90}, n_ff_poles 2, n_ff_piers 0 */
typedef struct JgSurrogateConfig {
    int n_pole_limits[3];
    float coverage_limits[3];
    int n_ff_poles;
    int n_ff_piers;
} JgSurrogateConfig;

/* generate_piers (geometry/fail_fast/piers.rs:25-226) as SPSurrogate::new calls it (sp_surrogate.rs:40-44): n_piers
 * segments inside the shape covering what the given poles (the first n_ff_poles surrogate poles) leave unrepresented.
 * verts_xy: the item's imported, counter-clockwise vertices (output of jg_item_import*); piers_xyxy: 4 * n_piers floats
 * (sx, sy, ex, ey), the form jg_item_upload takes. cfg == NULL or zero fields: the reference's release constants
 * (200 rays per angle, 90 angles, 100 grid points per dimension; its debug build uses 10 / 4 / 10). The reference's
 * panics (a ray end inside the shape, a NaN loss, no clipped ray) return JG_ERR_REFERENCE_PANIC. */
typedef struct JgPierConfig {
    int rays_per_angle;
    int n_angles;
    int n_points_per_dimension;
} JgPierConfig;
int jg_item_generate_piers(const float* verts_xy, int n_verts, const float* poles_xyr, int n_poles, int n_piers,
                           const JgPierConfig* cfg, float* piers_xyxy);

/* Importer::import_item for an ExtShape::SimplePolygon with shape modification disabled (io/import.rs:47-98,
 * 198-260) + Item::new's surrogate (entities/item.rs:28-53, fail_fast/sp_surrogate.rs:33-61, fail_fast/pole.rs:12-135):
 * cleans the vertex list (closing duplicate, approx-equal neighbours), orients it counter-clockwise, computes the
 * pole of inaccessibility, centres the shape on its centroid (pre_transform = (0, -cx, -cy)) and generates the
 * surrogate poles under the coverage limits. Outputs are what jg_item_upload takes. verts_xy must hold 2 * n_pts
 * floats; poles beyond poles_cap are counted in *n_poles but not written. Piers: jg_item_generate_piers. */
int jg_item_import(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float* verts_xy, int* n_verts,
                   float poi_xyr[3], float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2], float* area);

/* The same with Importer's poly_simpl_tolerance (io/import.rs:31-45, geometry/original_shape.rs:33-45): the centred shape
 * goes through simplify_shape in Inflate mode (geometry/shape_modification.rs:48-144) and SPolygon::new again before the
 * surrogate is generated; *area stays the original shape's. narrow_concavity_cutoff_ratio >= 0 (lbf default 0.01) adds
 * close_narrow_concavities (shape_modification.rs:374-502) and the second simplification with tolerance / 10; < 0 = None. */
int jg_item_import_simplified(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float simplify_tolerance,
                              float narrow_concavity_cutoff_ratio, float* verts_xy, int* n_verts, float poi_xyr[3],
                              float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2], float* area);
/* close_narrow_concavities (geometry/shape_modification.rs:374-502) on a clean counter-clockwise vertex list + SPolygon::new */
int jg_shape_close_concavities(const float* pts_xy, int n_pts, int mode, float max_distance_ratio, float* verts_xy, int* n_verts,
                               float poi_xyr[3]);
/* simplify_shape (geometry/shape_modification.rs:48-144) on a clean counter-clockwise vertex list; mode 0 Inflate (result
 * is a superset: items, holes, zones), 1 Deflate (subset: container boundaries); followed by SPolygon::new (poi). */
int jg_shape_simplify(const float* pts_xy, int n_pts, int mode, float max_area_change_ratio, float* verts_xy, int* n_verts,
                      float poi_xyr[3]);

/* ... with the whole ShapeModifyConfig (geometry/shape_modification.rs:22-46, original_shape.rs:23-51): offset > 0 inflates
 * the centred shape first (jg_shape_offset: the repo's own offsetter, not bit-equal to geo-buffer), then simplification /
 * concavity closing as in jg_item_import_simplified (negative = None). verts_xy must hold verts_cap >= 2 * n_pts vertices. */
int jg_item_import_modified(const float* pts_xy, int n_pts, const JgSurrogateConfig* scfg, float offset, float simplify_tolerance,
                            float narrow_concavity_cutoff_ratio, float* verts_xy, int verts_cap, int* n_verts, float poi_xyr[3],
                            float* poles_xyr, int poles_cap, int* n_poles, float pre_transform_xy[2], float* area);
/* SPolygon::new for a shape that is registered as it is (container outer boundary, hole, quality zone;
 * geometry/primitives/simple_polygon.rs:40-71): orientation normalised to counter-clockwise (written to
 * verts_ccw_xy when non-null, 2 * n_pts floats) and the pole of inaccessibility jg_cde_register_hazard wants. */
int jg_polygon_poi(const float* pts_xy, int n_pts, float poi_xyr[3], float* verts_ccw_xy);
/* offset_shape (geometry/shape_modification.rs:330-372; ShapeModifyConfig::offset = half the minimum separation): mode 0
 * Inflate (items, holes, zones), 1 Deflate (container boundaries), distance >= 0. The repo's own offsetter - miter joins,
 * corners beyond miter_limit * distance (default 2) cut on the tangents of the round join - NOT geo-buffer: an inflated
 * shape contains the reference's rounded one but is not bit-equal to it. pts: clean counter-clockwise vertices; the result
 * passes through the importer's cleaning and SPolygon::new (poi). At most 2 * n_pts vertices (verts_cap).
 * JG_ERR_REFERENCE_PANIC: the offset polygon intersects itself (separation too large for the shape). */
int jg_shape_offset(const float* pts_xy, int n_pts, int mode, float distance, float miter_limit, float* verts_xy, int verts_cap,
                    int* n_verts, float poi_xyr[3]);
/* import_simple_polygon + SPolygon::new for a container boundary, hole or quality zone (io/import.rs:100-195,
 * 198-211): vertex cleaning as in jg_item_import, then jg_polygon_poi. verts_xy must hold 2 * n_pts floats. */
int jg_shape_import(const float* pts_xy, int n_pts, float* verts_xy, int* n_verts, float poi_xyr[3]);

/* ------------------------------------------------------------------------------------------ */
/* In-process multi-device pool                                                                */
/* ------------------------------------------------------------------------------------------ */

/* One engine per CUDA device of this process (BASELINE north_star: "the query batch shards across
 * the GPUs of one box ... results are gathered on the host"). Layout, items and filters are REPLICATED on every device;
 * a host batch is cut into contiguous shards [g n/G, (g+1) n/G), device g copies its shard in, resolves it and copies the
 * results into ITS slice of the caller's one result array; there is no collective on the query path. One host thread per
 * device does that device's enqueueing. Batches smaller than min_shard queries per device use fewer devices. The *_async
 * forms return at once (segment tables are copied; `transforms` / `out` must stay valid - pinned - until jg_pool_sync). */
typedef struct JgPool JgPool;
int jg_pool_create(const int* devices, int n_devices, const JgConfig* cfg, JgPool** out);
void jg_pool_destroy(JgPool* p);
int jg_pool_size(const JgPool* p);
JgEngine* jg_pool_engine(JgPool* p, int k); /* the engine of the k-th device: knobs, stats, jg_layout_broadcast */
const char* jg_pool_last_error(const JgPool* p);
int jg_pool_set_min_shard(JgPool* p, int64_t min_queries_per_device);
/* The sharding rule as a pure function (no device needed; what the CPU tests check): device k of n_devices resolves queries
 * [lo, hi) of an n-query batch; devices the batch is not spread over get lo == hi == n. */
int jg_pool_shard_bounds(int64_t n, int k, int n_devices, int64_t min_shard, int64_t* lo, int64_t* hi);
int jg_pool_item_upload(JgPool* p, int item_id, const float* verts_xy, int n_verts, const float* poi_xy, const float* poles_xyr,
                        int n_poles, const float* piers, int n_piers);
int jg_pool_item_set_rotations(JgPool* p, int item_id, const float* rot_cos_sin, int n_rot);
int jg_pool_filter_upload(JgPool* p, int filter_id, const uint32_t* bits, int words);
int jg_pool_layout_upload(JgPool* p, const void* blob, size_t bytes, uint64_t epoch);
int jg_pool_cde_sync(JgCde* c, JgPool* p, uint64_t* epoch); /* flatten once, replicate */
int jg_pool_detect_multi(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int filter_id,
                         int mode, uint8_t* out);
int jg_pool_detect_multi_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                               int filter_id, int mode, uint8_t* out);
int jg_pool_detect_multi_indexed(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                                 const float* txy, int filter_id, int mode, uint8_t* out);
int jg_pool_detect_multi_indexed_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                                       const float* txy, int filter_id, int mode, uint8_t* out);
int jg_pool_detect_multi_lattice(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                                 const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out);
int jg_pool_detect_multi_lattice_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                                       const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out);
int jg_pool_collect_multi(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                          const uint32_t* seed_bits, uint32_t* out_bits, int words);
int jg_pool_collect_multi_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                                const uint32_t* seed_bits, uint32_t* out_bits, int words);
int jg_pool_sync(JgPool* p);

/* ------------------------------------------------------------------------------------------ */
/* Host mirror of CDEngine's update path                                                       */
/* ------------------------------------------------------------------------------------------ */

/* CDEngine::new (cd_engine.rs:34-57) with no hazards yet; static hazards are registered next with
 * dynamic = 0 (the Exterior must be among them before the first flatten). */
int jg_cde_create(const float bbox4[4], const JgConfig* cfg, JgCde** out);
void jg_cde_destroy(JgCde* c);
const char* jg_cde_last_error(const JgCde* c);

/* CDEngine::register_hazard (cd_engine.rs:59-69). shape = vertices (CCW) + poi centre; entity =
 * (kind, id, quality, pk). Returns the HazKey (slot index | version << 32) through *hkey. */
int jg_cde_register_hazard(JgCde* c, int kind, int64_t id, int64_t quality, uint64_t pk, const float* verts_xy,
                           int n_verts, const float* poi_xy, int dynamic, uint64_t* hkey);
/* Layout::place_item (entities/layout.rs:40-60) + PlacedItem::new (entities/placed_item.rs:27-36): the
 * item's collision shape is cloned under the composed transformation (cos, sin, tx, ty) - every vertex
 * and the poi centre by TRANSFORM_FORMULA (geometry/primitives/point.rs:26-31, three f32 roundings),
 * bbox rebuilt by a min/max fold - and registered as a dynamic HazardEntity::PlacedItem{id = item_id, pk}.
 * pk is the caller's PItemKey (slot index | version << 32 of its own placed-items SlotMap). */
int jg_cde_place_item(JgCde* c, int64_t item_id, uint64_t pk, const float* item_verts_xy, int n_verts,
                      const float* item_poi_xy, const float transform4[4], uint64_t* hkey);
/* Layout::remove_item (entities/layout.rs:62-74): deregisters the PlacedItem hazard carrying `pk` */
int jg_cde_remove_item(JgCde* c, uint64_t pk);
/* CDEngine::deregister_hazard_by_key (cd_engine.rs:87-96) */
int jg_cde_deregister_hazard(JgCde* c, uint64_t hkey);
/* CDEngine::deregister_hazard_by_entity (cd_engine.rs:72-85): the hazard whose HazardEntity equals (kind, id, quality,
 * pk); JG_ERR_INVALID when there is none (the reference panics: "Cannot deregister hazard that is not registered") */
int jg_cde_deregister_hazard_by_entity(JgCde* c, int kind, int64_t id, int64_t quality, uint64_t pk);
/* Layout::swap_container (entities/layout.rs:40-50): `base` plays Container::base_cde of the new container (its bbox,
 * config and static hazards); *c becomes a clone of it with every DYNAMIC hazard of the old *c re-registered in the
 * order of the old CDE's snapshot (cd_engine.rs:98-106: slot order). Snapshots taken before the swap are dropped. */
int jg_cde_swap_container(JgCde* c, const JgCde* base);
/* CDEngine::clone (what Layout::new does with Container::base_cde, entities/layout.rs:24-31) */
int jg_cde_clone(const JgCde* c, JgCde** out);
/* counts register / deregister calls: the epoch jg_cde_sync_to_engine stamps on the upload */
uint64_t jg_cde_epoch(const JgCde* c);
/* CDEngine::haz_key_from_pi_key (cd_engine.rs:333-341); *hkey = 0 when absent */
int jg_cde_hkey_from_pk(const JgCde* c, uint64_t pk, uint64_t* hkey);
/* CDEngine::save / restore (cd_engine.rs:98-149); snapshots are owned by the JgCde */
int jg_cde_save(JgCde* c, int* snapshot_id);
int jg_cde_restore(JgCde* c, int snapshot_id);
/* number of slots of the hazard SlotMap (collector / filter bitsets must cover it) */
int jg_cde_slot_capacity(const JgCde* c);
int jg_cde_n_hazards(const JgCde* c);
uint64_t jg_cde_hkey_exterior(const JgCde* c);
/* HazKeyFilter::from_irrelevant_qzones (filter.rs:21-41): bit set iff InferiorQualityZone with
 * quality < required_quality */
int jg_cde_filter_qzones(const JgCde* c, int required_quality, uint32_t* bits, int words);

/* Flattens quadtree + hazards into the SoA blob. Returns the byte size needed; writes it when
 * cap >= size. */
int64_t jg_cde_flatten(JgCde* c, void* buf, int64_t cap);
/* Brings the host mirror's incremental device image up to date: only the quadtree nodes and hazard slots that register /
 * deregister / restore touched since the previous call are re-serialised (what qt_node.rs:38-94 touches in the reference's
 * pointer tree). Returns the image (owned by the JgCde, valid until the next call), its lineage and sequence number, and
 * the byte ranges that changed (n_ranges = 0: a new lineage, take everything). Feed to jg_layout_patch. */
int jg_cde_image_sync(JgCde* c, const void** image, size_t* bytes, uint64_t* lineage, uint64_t* seq, const uint32_t** ranges_off_len, int* n_ranges);
/* jg_cde_image_sync + jg_layout_patch in one call (JG_INCREMENTAL=0: jg_cde_flatten + jg_layout_upload); returns the epoch */
int jg_cde_sync_to_engine(JgCde* c, JgEngine* e, uint64_t* epoch);
/* Canonical pre-order dump of the quadtree (u32 words) for tests: per node level, bbox bits,
 * has_children, n_hazards, n_active_edges, then entries sorted (presence desc, slot asc) with
 * ff_bbox and edges. Returns the word count; writes at most cap words. */
int64_t jg_cde_tree_dump(const JgCde* c, uint32_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* JAGUA_CUDA_H */
