// C ABI of the device engine (include/jagua_cuda.h): HBM residency of layout / items / filters,
// stream-overlapped batching of host buffers, kernel launches, counters.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/jagua_cuda.h"
#include "cde_host.hpp"
#include "launch.hpp"

using namespace jg;

#define JG_STREAM_MAX_CHUNKS 32          // copy pieces of a streamed batch (and arrival-counter slots)
#define JG_STREAM_CHUNK (1 << 16)        // ... of at least this many queries
#define JG_STREAM_MIN_QUERIES (1 << 17)  // smaller host batches keep the chunked pipeline

namespace {

struct ItemDev {
    float2* verts = nullptr;
    int n_verts = 0;
    float poi_x = 0, poi_y = 0;
    float4* poles = nullptr;
    int n_poles = 0;
    float4* piers = nullptr;
    int n_piers = 0;
    float2* rots = nullptr;  // (cos, sin) per allowed rotation (jg_item_set_rotations)
    int n_rots = 0;
    float2* rot_ext = nullptr;     // per allowed rotation: (x_max, y_max) of the rotated, untranslated shape (search rounds)
    std::vector<float> h_verts;    // host copy of the vertices (extents are computed on the host)
};
struct FilterDev {
    uint32_t* bits = nullptr;
    int words = 0;
    uint8_t* kinds = nullptr;  // node summary under this filter (k_node_kinds), valid for kinds_version
    size_t kinds_cap = 0;
    uint64_t kinds_version = 0;
    int grid_ok = 0;           // k_grid_build's verdict for this filter's summary (0: its batches take the walk kernels)
    // jg_filter_upload_keys: versioned HazKeys; `bits` is re-derived from them for every layout (bits_version)
    bool keyed = false;
    std::vector<uint64_t> keys;
    uint64_t bits_version = 0;
};
// One in-flight chunk of a host-pointer batch: its own stream and staging buffers.
struct ChunkCtx {
    cudaStream_t st = nullptr;
    float4* d_xf = nullptr;
    uint8_t* d_out = nullptr;  // verdict bytes or collector words
    size_t cap_q = 0, cap_out = 0;
    DetectScratch sc{nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap_sv = 0, cap_ct = 0;
    cudaEvent_t done = nullptr;  // join point of device-resident batches spread over the chunk streams
};

}  // namespace

struct JgEngine {
    int device = 0;
    JgConfig cfg{};
    int n_sms = 148;
    std::string err;

    // Two layout buffers: an upload fills the one that is not current on its own stream (st_up) while queries on
    // the current one are still running; queries launched afterwards wait for `ready` on the device, never on the host.
    struct LayoutBuf {
        void* d_blob = nullptr;
        size_t blob_cap = 0;
        uint8_t* d_kinds = nullptr;  // node summary under NoFilter (+ GridDev + GridCollect)
        size_t kinds_cap = 0;
        void* h_stage = nullptr;     // pinned copy of the blob (the caller's buffer may be pageable / short-lived)
        size_t stage_cap = 0;
        cudaEvent_t ready = nullptr; // copy + summaries of the last upload into this buffer are done
        std::vector<cudaEvent_t> marks;  // recorded on every query stream when this buffer became current
        bool marks_valid = false;
        int expect_grid = -1;        // host verdict to be confirmed by k_grid_build's (-1: nothing pending)
        // incremental images (jg_layout_patch): which image lineage this buffer holds (0: none) and the byte ranges that
        // changed since it was last written
        uint64_t img_lineage = 0;
        std::vector<ImageRange> pending;
        void* d_patch = nullptr;     // device copy of a patch buffer: {n, (off, len, data_off) x n, data}
        size_t patch_cap = 0;
    } bufs[2];
    int cur = 0;
    cudaStream_t st_up = nullptr;
    int* h_grid_ok = nullptr;  // pinned, [2]: k_grid_build's verdict of the last upload into each buffer
    int grid_ok_cur = 0;       // the grid premise holds for the current layout
    void* d_blob = nullptr;    // = bufs[cur].d_blob
    JgLayoutHeader hdr{};
    DevLayout L{};
    bool has_layout = false;
    uint64_t epoch = 0;
    uint64_t layout_version = 0;  // bumped by every upload; per-filter node summaries are keyed on it
    uint64_t img_lineage = 0, img_seq = 0;  // the incremental image last received (jg_layout_patch)
    uint64_t stat_patched = 0;
    std::vector<uint32_t> slot_versions;  // SlotMap version per hazard slot of the current layout (blob section haz_version)
    int* h_blob_bad = nullptr;    // pinned, [2]: k_blob_validate's count of violations in the last device-blob upload into each buffer
    int* d_blob_bad = nullptr;
    bool slot_versions_valid = false;
    // layout broadcast over NCCL (jg_engine_comm_init / jg_layout_broadcast): communicator, its own stream, the fixed-capacity
    // exchange buffer and the pinned read-back of {header, validation verdict}
    void* nccl_comm = nullptr;
    int comm_rank = -1, comm_world = 0;
    cudaStream_t st_bc = nullptr;
    void* d_bcast = nullptr;
    size_t bcast_cap = 0;
    void* h_bcast_stage = nullptr;  // pinned: root's copy of the blob on its way to d_bcast
    JgLayoutHeader* h_bcast_hdr = nullptr;  // pinned: header of the received blob (+ one int: violations found by the device check)
    cudaEvent_t ev_bc_consumed = nullptr;   // the last blob has been copied out of d_bcast
    void* d_hdr_scratch = nullptr;          // device copy of the header k_bcast_check parsed
    DetectPlan plan{0, 0, 0, 0, 1 << 21};
    long long split_min = -1;  // jg_engine_set_split_min override (-1: the default of make_detect_plan)
    bool force_walk = false;  // jg_engine_set_detect_path: keep the quadtree-walk kernels even when the grid applies
    bool use_fused = false;   // JG_PATH_FUSED: one fused, copy-following kernel per batch (experimental, slower)
    uint8_t* d_kinds = nullptr;   // = bufs[cur].d_kinds

    std::map<int, ItemDev> items;
    std::map<int, FilterDev> filters;

    cudaStream_t compute = nullptr;
    ChunkCtx main_ctx;  // scratch of the *_device calls (compute stream)
    std::vector<ChunkCtx> chunks;
    // streamed host batches (k_detect_fused follows the copy front): whole-batch buffers on `compute`, the
    // copies on `st_copy`, the arrival counter values in pinned host memory
    ChunkCtx stream_ctx;
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_reset = nullptr;
    int* h_avail = nullptr;
    bool stream_pending = false;  // a streamed batch may still be in flight: check its error flag at the next sync
    int* h_err = nullptr;
    int* h_gate = nullptr;  // jg_timer_start_gated: the flag the gate kernel waits for (pinned, read over PCIe)
    bool gate_armed = false;
    int64_t chunk_queries = 1 << 18;
    bool auto_schedule = true;  // piece sizes chosen per batch (detect_multi_host); jg_engine_set_batching pins them
    uint64_t next_chunk = 0;
    int detect_group = 0, collect_group = 0;  // 0 = auto

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // jg_timer_*
    bool timing_valid = false;
    uint32_t* d_flush = nullptr;
    size_t flush_words = 0;

    // fences: completion points of everything enqueued so far (one event per query stream), a small ring
    struct Fence {
        std::vector<cudaEvent_t> ev;
        uint64_t id = 0;
    };
    std::vector<Fence> fences;
    uint64_t next_fence = 1;
    unsigned long long* d_search = nullptr;  // search rounds: {min key, n_free, best txy bits, best rot idx}
    unsigned long long* h_search = nullptr;  // pinned copy
    int64_t search_last_n = 0;
    uint64_t stat_launches = 0, stat_queries = 0, stat_h2d = 0, stat_d2h = 0, stat_uploads = 0;
    // JG_TRACE=1: device timeline of the pieces of the last host batch, printed by jg_sync (debugging aid)
    std::vector<cudaEvent_t> trace_ev;
    std::vector<int64_t> trace_m;
    size_t trace_used = 0;
    double trace_host_t0 = 0.0;
};
static bool trace_on() {
    static const bool on = getenv("JG_TRACE") != nullptr;
    return on;
}
static cudaEvent_t trace_mark(JgEngine* e, cudaStream_t st) {
    if (e->trace_used == e->trace_ev.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->trace_ev.push_back(ev);
    }
    cudaEvent_t ev = e->trace_ev[e->trace_used++];
    cudaEventRecord(ev, st);
    return ev;
}

#define JG_CUDA(e, call)                                                                             \
    do {                                                                                             \
        cudaError_t _err = (call);                                                                   \
        if (_err != cudaSuccess) {                                                                   \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_err);                         \
            return JG_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

static int fail(JgEngine* e, int code, const char* msg) {
    e->err = msg;
    return code;
}

static int auto_group(int n_verts) {
    // lanes per query in stage 2 / collect: enough lanes to cover the polygon in ~3 rounds
    if (n_verts <= 12) return 4;
    if (n_verts <= 48) return 8;
    if (n_verts <= 128) return 16;
    return 32;
}

static int ensure_ctx(JgEngine* e, ChunkCtx& c, size_t n_q, size_t out_bytes, bool own_xf) {
    if (own_xf && c.cap_q < n_q) {
        if (c.d_xf) cudaFree(c.d_xf);
        JG_CUDA(e, cudaMalloc(&c.d_xf, n_q * sizeof(float4)));
        c.cap_q = n_q;
    }
    if (out_bytes && c.cap_out < out_bytes) {
        if (c.d_out) cudaFree(c.d_out);
        JG_CUDA(e, cudaMalloc(&c.d_out, out_bytes));
        c.cap_out = out_bytes;
    }
    if (c.cap_sv < n_q) {
        cudaFree(c.sc.sv_entry);
        c.sc.sv_entry = nullptr;
        c.cap_sv = 0;
        if (!c.sc.ctl) {
            JG_CUDA(e, cudaMalloc(&c.sc.ctl, sizeof(DetectCtl)));
            JG_CUDA(e, cudaMemset(c.sc.ctl, 0, sizeof(DetectCtl)));
        }
        JG_CUDA(e, cudaMalloc(&c.sc.sv_entry, n_q * sizeof(int)));
        c.cap_sv = n_q;
    }
    // the hand-over list of the containment kernel: only slots that can see a launch large enough to use it (48 B/query)
    if ((long long)c.cap_sv >= e->plan.split_min && c.cap_ct < c.cap_sv) {
        cudaFree(c.sc.ct_meta);
        cudaFree(c.sc.ct_bb);
        cudaFree(c.sc.ct_xf);
        c.sc.ct_meta = nullptr, c.sc.ct_bb = nullptr, c.sc.ct_xf = nullptr;
        c.cap_ct = 0;
        JG_CUDA(e, cudaMalloc(&c.sc.ct_meta, c.cap_sv * sizeof(int4)));
        JG_CUDA(e, cudaMalloc(&c.sc.ct_bb, c.cap_sv * sizeof(float4)));
        JG_CUDA(e, cudaMalloc(&c.sc.ct_xf, c.cap_sv * sizeof(float4)));
        c.cap_ct = c.cap_sv;
    }
    return JG_OK;
}

static void free_ctx(ChunkCtx& c) {
    cudaFree(c.d_xf), cudaFree(c.d_out);
    cudaFree(c.sc.ctl), cudaFree(c.sc.sv_entry), cudaFree(c.sc.ct_meta), cudaFree(c.sc.ct_bb), cudaFree(c.sc.ct_xf);
    if (c.done) cudaEventDestroy(c.done);
    c = ChunkCtx{};
}

// Structural checks of a blob that arrived in device memory (jg_layout_upload_device, jg_layout_broadcast): the twin of the
// host loop in upload_common. The kernels index with these fields unchecked, so a malformed blob must not be served.
__global__ void k_blob_validate(const DevLayout L, int n_entries, int n_edges, int n_haz_verts, int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool b = false;
    if (i < L.n_nodes) {
        const int4 m = L.node_meta[i];
        const int n_ent = (m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF);
        if (m.x >= 0 && ((m.x & 3) != 1 || m.x + 4 > L.n_nodes)) b = true;
        if (m.y < 0 || m.y + n_ent > n_entries) b = true;
        if (m.w >= L.n_nodes || (i > 0 && m.w < 0 && m.w != -2)) b = true;
        const float4 r = L.node_bbox[i];
        if (!(fabsf(r.x) < 1e30f && fabsf(r.y) < 1e30f && fabsf(r.z) < 1e30f && fabsf(r.w) < 1e30f && r.x <= r.z && r.y <= r.w)) b = true;
    }
    if (i < n_entries) {
        const int4 em = L.entry_meta[i];
        if (em.x < 0 || em.x >= L.n_slots || em.y < 0 || em.z < 0 || em.y + em.z > n_edges) b = true;
    }
    if (i < L.n_slots) {
        const int4 hm = L.haz_meta[i];
        if (hm.y < 0 || hm.z < 0 || hm.y + hm.z > n_haz_verts) b = true;
    }
    if (i == 0 && (L.exterior_slot < 0 || L.exterior_slot >= L.n_slots)) b = true;
    if (b) atomicAdd(bad, 1);
}
static cudaError_t launch_blob_validate(const DevLayout& L, const JgLayoutHeader& h, int* d_bad, cudaStream_t st) {
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    uint32_t n = h.n_nodes > h.n_entries ? h.n_nodes : h.n_entries;
    if (h.n_slots > n) n = h.n_slots;
    k_blob_validate<<<(n + 255) / 256, 256, 0, st>>>(L, (int)h.n_entries, (int)h.n_edges, (int)h.n_haz_verts, d_bad);
    return cudaGetLastError();
}

extern "C" {

int jg_engine_create(int device, const JgConfig* cfg, JgEngine** out) {
    if (!cfg || !out) return JG_ERR_INVALID;
    *out = nullptr;
    if (cfg->quadtree_depth < 0 || cfg->quadtree_depth > JG_MAX_DEPTH || cfg->cd_threshold < 0 ||
        cfg->n_ff_poles < 0 || cfg->n_ff_piers < 0)
        return JG_ERR_INVALID;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return JG_ERR_CUDA;
    JgEngine* e = new JgEngine();
    e->device = device;
    e->cfg = *cfg;
    cudaError_t err = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (err == cudaSuccess) err = cudaGetDeviceProperties(&prop, device);
    if (err == cudaSuccess) {
        e->n_sms = prop.multiProcessorCount;
        err = cudaStreamCreateWithFlags(&e->compute, cudaStreamNonBlocking);
    }
    if (err == cudaSuccess) err = cudaEventCreate(&e->ev0);
    if (err == cudaSuccess) err = cudaEventCreate(&e->ev1);
    if (err == cudaSuccess) err = cudaEventCreate(&e->tm0);
    if (err == cudaSuccess) err = cudaEventCreate(&e->tm1);
    if (err != cudaSuccess) {
        delete e;
        return JG_ERR_CUDA;
    }
    e->main_ctx.st = e->compute;
    e->stream_ctx.st = e->compute;
    if (cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_reset, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc((void**)&e->h_avail, (JG_STREAM_MAX_CHUNKS + 2) * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        delete e;
        return JG_ERR_CUDA;
    }
    e->h_err = e->h_avail + JG_STREAM_MAX_CHUNKS;
    *e->h_err = 0;
    e->h_gate = e->h_avail + JG_STREAM_MAX_CHUNKS + 1;
    *e->h_gate = 1;
    if (cudaStreamCreateWithFlags(&e->st_up, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void**)&e->h_grid_ok, 4 * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->bufs[0].ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->bufs[1].ready, cudaEventDisableTiming) != cudaSuccess) {
        delete e;
        return JG_ERR_CUDA;
    }
    e->h_grid_ok[0] = e->h_grid_ok[1] = 0;
    e->h_blob_bad = e->h_grid_ok + 2;
    e->h_blob_bad[0] = e->h_blob_bad[1] = 0;
    if (cudaMalloc((void**)&e->d_blob_bad, 2 * sizeof(int)) != cudaSuccess) {
        delete e;
        return JG_ERR_CUDA;
    }
    e->chunks.resize(4);  // in-flight pieces of host batches (tools/pipeline_probe.py: 4 x 256 Ki is the best default)
    for (ChunkCtx& c : e->chunks)
        if (cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking) != cudaSuccess) {
            delete e;
            return JG_ERR_CUDA;
        }
    *out = e;
    return JG_OK;
}

void jg_engine_destroy(JgEngine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto& kv : e->items) {
        cudaFree(kv.second.verts);
        cudaFree(kv.second.poles);
        cudaFree(kv.second.piers);
        cudaFree(kv.second.rots);
        cudaFree(kv.second.rot_ext);
    }
    for (auto& kv : e->filters) cudaFree(kv.second.bits), cudaFree(kv.second.kinds);
    {
        cudaStream_t keep = e->stream_ctx.st;
        free_ctx(e->stream_ctx);
        (void)keep;  // = e->compute, destroyed below
    }
    if (e->st_copy) cudaStreamDestroy(e->st_copy);
    if (e->ev_reset) cudaEventDestroy(e->ev_reset);
    if (e->h_avail) cudaFreeHost(e->h_avail);
    for (JgEngine::LayoutBuf& B : e->bufs) {
        cudaFree(B.d_blob), cudaFree(B.d_kinds), cudaFree(B.d_patch);
        if (B.h_stage) cudaFreeHost(B.h_stage);
        if (B.ready) cudaEventDestroy(B.ready);
        for (cudaEvent_t ev : B.marks) cudaEventDestroy(ev);
    }
    for (JgEngine::Fence& f : e->fences)
        for (cudaEvent_t ev : f.ev) cudaEventDestroy(ev);
    if (e->d_search) cudaFree(e->d_search);
    if (e->h_search) cudaFreeHost(e->h_search);
    if (e->st_up) cudaStreamDestroy(e->st_up);
    if (e->h_grid_ok) cudaFreeHost(e->h_grid_ok);
    if (e->d_blob_bad) cudaFree(e->d_blob_bad);
    jg_engine_comm_destroy(e);
    if (e->d_hdr_scratch) cudaFree(e->d_hdr_scratch);
    for (ChunkCtx& c : e->chunks) {
        cudaStream_t st = c.st;
        free_ctx(c);
        if (st) cudaStreamDestroy(st);
    }
    free_ctx(e->main_ctx);
    if (e->d_flush) cudaFree(e->d_flush);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->tm0) cudaEventDestroy(e->tm0);
    if (e->tm1) cudaEventDestroy(e->tm1);
    if (e->compute) cudaStreamDestroy(e->compute);
    delete e;
}

const char* jg_last_error(const JgEngine* e) { return e ? e->err.c_str() : "null engine"; }

int jg_sync(JgEngine* e) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    for (ChunkCtx& c : e->chunks) JG_CUDA(e, cudaStreamSynchronize(c.st));
    if (e->st_copy) JG_CUDA(e, cudaStreamSynchronize(e->st_copy));
    if (e->st_up) JG_CUDA(e, cudaStreamSynchronize(e->st_up));
    JG_CUDA(e, cudaStreamSynchronize(e->compute));
    for (int b = 0; b < 2; b++)  // stream-ordered uploads: k_grid_build must have confirmed the host's verdict
        if (e->bufs[b].expect_grid >= 0) {
            const int want = e->bufs[b].expect_grid;
            e->bufs[b].expect_grid = -1;
            if (e->h_grid_ok[b] != want) return fail(e, JG_ERR_INVALID, "layout upload: host and device disagree on the grid premise");
        }

    if (trace_on() && e->trace_used >= 5) {
        fprintf(stderr, "[jg trace] piece: queries | t(start) t(h2d done) t(kernels done) t(d2h done)  [us since first mark]\n");
        for (size_t i = 1; i + 3 < e->trace_used + 0 && (i - 1) / 4 < e->trace_m.size(); i += 4) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, e->trace_ev[0], e->trace_ev[i]);
            cudaEventElapsedTime(&b, e->trace_ev[0], e->trace_ev[i + 1]);
            cudaEventElapsedTime(&c, e->trace_ev[0], e->trace_ev[i + 2]);
            cudaEventElapsedTime(&d, e->trace_ev[0], e->trace_ev[i + 3]);
            fprintf(stderr, "[jg trace] %8lld | %8.1f %8.1f %8.1f %8.1f\n", (long long)e->trace_m[(i - 1) / 4], a * 1e3, b * 1e3, c * 1e3, d * 1e3);
        }
        e->trace_used = 0;
        e->trace_m.clear();
    }
    if (e->stream_pending) {
        e->stream_pending = false;
        if (*e->h_err) {
            *e->h_err = 0;
            return fail(e, JG_ERR_CUDA, "streamed detect batch: a warp gave up waiting for its queries (copy lost or stalled)");
        }
    }
    return JG_OK;
}

static int validate_header(JgEngine* e, const JgLayoutHeader& h, size_t bytes) {
    if (bytes < sizeof(JgLayoutHeader) || h.magic != JG_LAYOUT_MAGIC || h.version != JG_LAYOUT_VERSION)
        return fail(e, JG_ERR_INVALID, "layout blob: bad magic / version");
    if (h.total_bytes > bytes || h.n_nodes == 0) return fail(e, JG_ERR_INVALID, "layout blob: truncated");
    if (h.root_level > JG_MAX_DEPTH) return fail(e, JG_ERR_INVALID, "layout blob: quadtree too deep");
    if (h.n_nodes > JG_MAX_NODES) return fail(e, JG_ERR_CAPACITY, "layout blob: more than 2^20 quadtree nodes");
    struct {
        uint32_t off, n, sz;
    } secs[] = {{h.off_node_meta, h.n_nodes, 16},  {h.off_node_bbox, h.n_nodes, 16},    {h.off_node_nactive, h.n_nodes, 4},
                {h.off_entry_meta, h.n_entries, 16}, {h.off_entry_ffbbox, h.n_entries, 16}, {h.off_edges, h.n_edges, 16},
                {h.off_haz_meta, h.n_slots, 16},   {h.off_haz_bbox, h.n_slots, 16},     {h.off_haz_poi, h.n_slots, 16},
                {h.off_haz_verts, h.n_haz_verts, 8},  {h.off_haz_version, h.n_slots, 4}};
    for (auto& s : secs)
        if ((s.off & 15u) || (uint64_t)s.off + (uint64_t)s.n * s.sz > h.total_bytes)
            return fail(e, JG_ERR_INVALID, "layout blob: section out of bounds / misaligned");
    return JG_OK;
}

static DevLayout dev_layout_of(const void* d_blob, const JgLayoutHeader& h) {
    const char* b = (const char*)d_blob;
    DevLayout L;
    L.node_meta = (const int4*)(b + h.off_node_meta);
    L.node_bbox = (const float4*)(b + h.off_node_bbox);
    L.node_nactive = (const uint32_t*)(b + h.off_node_nactive);
    L.entry_meta = (const int4*)(b + h.off_entry_meta);
    L.entry_ff = (const float4*)(b + h.off_entry_ffbbox);
    L.edges = (const float4*)(b + h.off_edges);
    L.haz_meta = (const int4*)(b + h.off_haz_meta);
    L.haz_bbox = (const float4*)(b + h.off_haz_bbox);
    L.haz_poi = (const float4*)(b + h.off_haz_poi);
    L.haz_verts = (const float2*)(b + h.off_haz_verts);
    L.cde_bbox = Rect{h.cde_bbox[0], h.cde_bbox[1], h.cde_bbox[2], h.cde_bbox[3]};
    L.n_nodes = (int)h.n_nodes;
    L.n_slots = (int)h.n_slots;
    L.cd_threshold = (int)h.cd_threshold;
    L.exterior_slot = (int)h.exterior_slot;
    L.n_entries = (int)h.n_entries;
    L.n_edges = (int)h.n_edges;
    return L;
}
static void bind_layout(JgEngine* e, const JgLayoutHeader& h) {
    e->hdr = h;
    e->L = dev_layout_of(e->d_blob, h);
    e->has_layout = true;
}

// Scatter of a patch buffer into a layout image: patch = {n, (byte offset, byte length, data offset) x n, data}; all
// multiples of 4. One CTA per range.
__global__ void k_apply_patch(const uint32_t* __restrict__ patch, unsigned char* __restrict__ blob) {
    const uint32_t off = patch[1 + 3 * blockIdx.x], len = patch[2 + 3 * blockIdx.x], src = patch[3 + 3 * blockIdx.x];
    const uint32_t* s = (const uint32_t*)((const unsigned char*)patch + src);
    uint32_t* d = (uint32_t*)(blob + off);
    for (uint32_t i = threadIdx.x; i < len / 4; i += blockDim.x) d[i] = s[i];
}
struct PatchInfo {
    uint64_t lineage, seq;
    const ImageRange* ranges;
    int n_ranges;
};

// trusted_hdr != null: a device blob whose header is already on the host and whose structure the device has checked
// (jg_layout_broadcast): no read-back, no second check, and the copy out of `src` is signalled by ev_bc_consumed
// instead of a host wait.
// patch != null: `src` is the current state of an incremental image of the library's own host mirror (position-stable, trusted:
// no structural re-check); only the ranges that changed since the target buffer was last written are copied.
static int upload_common(JgEngine* e, const void* src, size_t bytes, uint64_t epoch, bool src_on_device, const JgLayoutHeader* trusted_hdr = nullptr,
                         const PatchInfo* patch = nullptr) {
    if (!e || !src) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JgLayoutHeader h;
    if (bytes < sizeof(h)) return fail(e, JG_ERR_INVALID, "layout blob: truncated");
    if (trusted_hdr) h = *trusted_hdr;
    else if (src_on_device) JG_CUDA(e, cudaMemcpy(&h, src, sizeof(h), cudaMemcpyDeviceToHost));
    else std::memcpy(&h, src, sizeof(h));
    int rc = validate_header(e, h, bytes);
    if (rc != JG_OK) return rc;
    if (!src_on_device && !patch) {
        // structural checks the kernels rely on: children blocks at 1 + 4k, indices in range
        const int32_t* nm = (const int32_t*)((const char*)src + h.off_node_meta);
        for (uint32_t i = 0; i < h.n_nodes; i++) {
            const int32_t cb = nm[4 * i], eb = nm[4 * i + 1], cnt = nm[4 * i + 2], par = nm[4 * i + 3];
            const uint32_t n_ent = (uint32_t)(cnt & 0xFFFF) + (uint32_t)((cnt >> 16) & 0xFFFF);
            if (cb >= 0 && ((cb & 3) != 1 || (uint32_t)cb + 4 > h.n_nodes)) return fail(e, JG_ERR_INVALID, "layout blob: bad child block");
            if (eb < 0 || (uint32_t)eb + n_ent > h.n_entries) return fail(e, JG_ERR_INVALID, "layout blob: bad entry range");
            if (par >= (int32_t)h.n_nodes || (i > 0 && par < 0 && par != -2)) return fail(e, JG_ERR_INVALID, "layout blob: bad parent");
        }
        // the shortcut descents of detect.cu assume finite, non-empty node boxes
        const float* nb = (const float*)((const char*)src + h.off_node_bbox);
        for (uint32_t i = 0; i < h.n_nodes; i++) {
            const float x0 = nb[4 * i], y0 = nb[4 * i + 1], x1 = nb[4 * i + 2], y1 = nb[4 * i + 3];
            if (!(std::fabs(x0) < 1e30f && std::fabs(y0) < 1e30f && std::fabs(x1) < 1e30f && std::fabs(y1) < 1e30f && x0 <= x1 && y0 <= y1))
                return fail(e, JG_ERR_INVALID, "layout blob: node bbox not finite");
        }
        const int32_t* em = (const int32_t*)((const char*)src + h.off_entry_meta);
        for (uint32_t i = 0; i < h.n_entries; i++) {
            if (em[4 * i] < 0 || (uint32_t)em[4 * i] >= h.n_slots) return fail(e, JG_ERR_INVALID, "layout blob: bad hazard slot");
            if (em[4 * i + 1] < 0 || em[4 * i + 2] < 0 || (uint32_t)em[4 * i + 1] + (uint32_t)em[4 * i + 2] > h.n_edges)
                return fail(e, JG_ERR_INVALID, "layout blob: bad edge range");
        }
    }
    if (src_on_device && !trusted_hdr) {
        // A blob that arrived in device memory is checked there, BEFORE anything reads it: one small kernel on the upload
        // stream and a 4-byte read-back (the engine keeps serving the previous layout meanwhile, and keeps it on failure).
        JG_CUDA(e, launch_blob_validate(dev_layout_of(src, h), h, e->d_blob_bad, e->st_up));
        JG_CUDA(e, cudaMemcpyAsync(e->h_blob_bad, e->d_blob_bad, sizeof(int), cudaMemcpyDeviceToHost, e->st_up));
        JG_CUDA(e, cudaStreamSynchronize(e->st_up));
        e->stat_launches += 1;
        if (e->h_blob_bad[0])
            return fail(e, JG_ERR_INVALID, "layout upload: malformed device blob (index out of range / bad child block / non-finite box)");
    }
    // Which kernels serve this layout (grid or walk) is decided on the host: by the producer (flags) or here.
    int host_grid = -1;  // -1 unknown (device blob of a foreign producer): read k_grid_build's verdict back
    if (h.flags & JG_LAYOUT_FLAG_GRID_CHECKED) host_grid = (h.flags & JG_LAYOUT_FLAG_GRID_OK) ? 1 : 0;
    else if (!src_on_device) host_grid = blob_grid_ok((const uint8_t*)src) ? 1 : 0;
    // incremental images: which buffers can be patched (they hold this lineage, and no sync was missed)
    if (patch) {
        const bool cont = e->img_lineage == patch->lineage && e->img_seq + 1 == patch->seq;
        for (JgEngine::LayoutBuf& X : e->bufs) {
            if (!cont || X.img_lineage != patch->lineage) {
                X.img_lineage = 0;
                X.pending.clear();
            } else {
                X.pending.insert(X.pending.end(), patch->ranges, patch->ranges + patch->n_ranges);
            }
        }
        e->img_lineage = patch->lineage, e->img_seq = patch->seq;
    } else {
        e->img_lineage = 0;
        for (JgEngine::LayoutBuf& X : e->bufs) X.img_lineage = 0, X.pending.clear();
    }
    static const bool force_sync = getenv("JG_SYNC_UPLOAD") != nullptr;
    // stream-ordered path: needs a current layout to keep serving, a host verdict, and no per-filter summaries
    // (those are rebuilt lazily on the compute stream and are not double-buffered)
    const bool fast = !force_sync && e->has_layout && host_grid >= 0 && e->filters.empty();
    const int nb = fast ? 1 - e->cur : e->cur;
    JgEngine::LayoutBuf& B = e->bufs[nb];
    const size_t kinds_need = detect_kinds_bytes((int)h.n_nodes);
    const bool grow = B.blob_cap < h.total_bytes || B.kinds_cap < kinds_need || (fast && !src_on_device && B.stage_cap < h.total_bytes);
    if (!fast || grow) {
        rc = jg_sync(e);  // queries in flight finish on the old layout
        if (rc != JG_OK) return rc;
    }
    if (B.blob_cap < h.total_bytes) {
        if (B.d_blob) cudaFree(B.d_blob);
        B.d_blob = nullptr;
        B.blob_cap = 0;
        const size_t cap = (size_t)h.total_bytes * 2 + 4096;
        JG_CUDA(e, cudaMalloc(&B.d_blob, cap));
        B.blob_cap = cap;
    }
    if (B.kinds_cap < kinds_need) {
        if (B.d_kinds) cudaFree(B.d_kinds);
        B.d_kinds = nullptr;
        B.kinds_cap = 0;
        JG_CUDA(e, cudaMalloc(&B.d_kinds, kinds_need * 2));
        B.kinds_cap = kinds_need * 2;
    }
    if (fast) {
        std::vector<cudaStream_t> qs{e->compute, e->st_copy};
        for (ChunkCtx& c : e->chunks) qs.push_back(c.st);
        // readers of buffer nb were launched before the other buffer became current: its marks cover them
        JgEngine::LayoutBuf& O = e->bufs[1 - nb];
        if (O.marks_valid)
            for (cudaEvent_t ev : O.marks) JG_CUDA(e, cudaStreamWaitEvent(e->st_up, ev, 0));
        else
            for (cudaStream_t st : qs) JG_CUDA(e, cudaStreamSynchronize(st));
        if (!src_on_device) {
            if (B.stage_cap < h.total_bytes) {
                if (B.h_stage) cudaFreeHost(B.h_stage);
                B.h_stage = nullptr;
                B.stage_cap = 0;
                JG_CUDA(e, cudaHostAlloc(&B.h_stage, (size_t)h.total_bytes * 2 + 4096, cudaHostAllocDefault));
                B.stage_cap = (size_t)h.total_bytes * 2 + 4096;
            }
            JG_CUDA(e, cudaEventSynchronize(B.ready));  // the previous copy out of this staging buffer (two uploads ago)
            size_t patch_data = 0;
            if (patch && B.img_lineage == patch->lineage)
                for (const ImageRange& r : B.pending) patch_data += r.len;
            const size_t patch_bytes = 4 + 12 * B.pending.size() + patch_data + 16;
            if (patch && B.img_lineage == patch->lineage && !B.pending.empty() && patch_bytes < h.total_bytes / 2 && patch_bytes <= B.stage_cap) {
                // ONE copy of {range table, range data} + a scatter kernel (a copy per range would cost more than the whole blob)
                if (B.patch_cap < B.stage_cap) {
                    cudaFree(B.d_patch);
                    B.d_patch = nullptr;
                    B.patch_cap = 0;
                    JG_CUDA(e, cudaMalloc(&B.d_patch, B.stage_cap));
                    B.patch_cap = B.stage_cap;
                }
                uint32_t* tab = (uint32_t*)B.h_stage;
                const uint32_t n_r = (uint32_t)B.pending.size();
                tab[0] = n_r;
                uint32_t at = (4 + 12 * n_r + 15) & ~15u;
                for (uint32_t k = 0; k < n_r; k++) {
                    const ImageRange& r = B.pending[k];
                    tab[1 + 3 * k] = r.off, tab[2 + 3 * k] = r.len, tab[3 + 3 * k] = at;
                    std::memcpy((char*)B.h_stage + at, (const char*)src + r.off, r.len);
                    at += r.len;
                }
                JG_CUDA(e, cudaMemcpyAsync(B.d_patch, B.h_stage, at, cudaMemcpyHostToDevice, e->st_up));
                k_apply_patch<<<n_r, 128, 0, e->st_up>>>((const uint32_t*)B.d_patch, (unsigned char*)B.d_blob);
                JG_CUDA(e, cudaGetLastError());
                e->stat_h2d += at;
                e->stat_patched++;
                e->stat_launches++;
            } else {
                std::memcpy(B.h_stage, src, h.total_bytes);
                JG_CUDA(e, cudaMemcpyAsync(B.d_blob, B.h_stage, h.total_bytes, cudaMemcpyHostToDevice, e->st_up));
                e->stat_h2d += h.total_bytes;
            }
            B.img_lineage = patch ? patch->lineage : 0;
            B.pending.clear();
        } else {
            // (a broadcast blob was read back and checked on st_bc, which the host has waited for: it is complete)
            JG_CUDA(e, cudaMemcpyAsync(B.d_blob, src, h.total_bytes, cudaMemcpyDeviceToDevice, e->st_up));
        }
        e->cur = nb;
        e->d_blob = B.d_blob;
        e->d_kinds = B.d_kinds;
        bind_layout(e, h);
        e->layout_version++;
        e->plan = make_detect_plan((int)h.n_nodes, (int)h.root_level);
        if (e->split_min >= 0) e->plan.split_min = e->split_min;
        JG_CUDA(e, launch_node_kinds(e->L, (int)h.root_level, nullptr, B.d_kinds, e->st_up));
        JG_CUDA(e, cudaMemcpyAsync(&e->h_grid_ok[nb], detect_grid_ok_ptr(B.d_kinds, (int)h.n_nodes), sizeof(int), cudaMemcpyDeviceToHost, e->st_up));
        JG_CUDA(e, cudaEventRecord(B.ready, e->st_up));
        B.expect_grid = host_grid;
        e->grid_ok_cur = host_grid;
        e->plan.use_grid = (host_grid && !e->force_walk) ? 1 : 0;
        // queries launched from now on read the new buffer: they wait for it on the device; the marks taken now
        // cover every reader of the other buffer
        if (B.marks.size() != qs.size()) {
            for (cudaEvent_t ev : B.marks) cudaEventDestroy(ev);
            B.marks.assign(qs.size(), nullptr);
            for (cudaEvent_t& ev : B.marks) JG_CUDA(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        }
        for (size_t i = 0; i < qs.size(); i++) {
            JG_CUDA(e, cudaEventRecord(B.marks[i], qs[i]));
            JG_CUDA(e, cudaStreamWaitEvent(qs[i], B.ready, 0));
        }
        B.marks_valid = true;
        if (trusted_hdr) JG_CUDA(e, cudaEventRecord(e->ev_bc_consumed, e->st_up));  // the next broadcast waits for this on the device
        else if (src_on_device) JG_CUDA(e, cudaStreamSynchronize(e->st_up));         // the caller may reuse its buffer on return
    } else {
        if (trusted_hdr) JG_CUDA(e, cudaStreamSynchronize(e->st_bc));  // (this path does not order `compute` behind the broadcast stream)
        JG_CUDA(e, cudaMemcpyAsync(B.d_blob, src, h.total_bytes,
                                   src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->compute));
        JG_CUDA(e, cudaStreamSynchronize(e->compute));
        if (!src_on_device) e->stat_h2d += h.total_bytes;
        B.img_lineage = patch ? patch->lineage : 0;
        B.pending.clear();
        if (trusted_hdr) JG_CUDA(e, cudaEventRecord(e->ev_bc_consumed, e->compute));
        e->cur = nb;
        e->d_blob = B.d_blob;
        e->d_kinds = B.d_kinds;
        bind_layout(e, h);
        e->layout_version++;
        e->plan = make_detect_plan((int)h.n_nodes, (int)h.root_level);
        if (e->split_min >= 0) e->plan.split_min = e->split_min;
        JG_CUDA(e, launch_node_kinds(e->L, (int)h.root_level, nullptr, B.d_kinds, e->compute));
        int ok = 0;
        JG_CUDA(e, detect_grid_usable(B.d_kinds, (int)h.n_nodes, e->compute, &ok));
        if (host_grid >= 0 && host_grid != ok)
            return fail(e, JG_ERR_INVALID, "layout upload: host and device disagree on the grid premise");
        e->grid_ok_cur = ok;
        e->plan.use_grid = (ok && !e->force_walk) ? 1 : 0;
        B.expect_grid = -1;
        e->bufs[0].marks_valid = e->bufs[1].marks_valid = false;  // everything was synchronised
    }
    e->stat_launches += 2;
    e->epoch = epoch;
    e->stat_uploads++;
    if (!src_on_device) {
        const uint32_t* hv = (const uint32_t*)((const char*)src + h.off_haz_version);
        e->slot_versions.assign(hv, hv + h.n_slots);
        e->slot_versions_valid = true;
    } else {
        e->slot_versions_valid = false;  // fetched from the device when a keyed filter needs them (find_filter)
    }
    return JG_OK;
}

int jg_layout_upload(JgEngine* e, const void* blob, size_t bytes, uint64_t epoch) {
    return upload_common(e, blob, bytes, epoch, false);
}
int jg_layout_upload_device(JgEngine* e, const void* d_blob, size_t bytes, uint64_t epoch) {
    return upload_common(e, d_blob, bytes, epoch, true);
}
int jg_layout_patch(JgEngine* e, const void* image, size_t bytes, uint64_t lineage, uint64_t seq, const uint32_t* ranges_off_len, int n_ranges,
                    uint64_t epoch) {
    if (!e || !image || lineage == 0 || n_ranges < 0 || (n_ranges > 0 && !ranges_off_len)) return JG_ERR_INVALID;
    static_assert(sizeof(ImageRange) == 8, "ranges cross the C ABI as (offset, length) pairs of u32");
    const PatchInfo pi{lineage, seq, (const ImageRange*)ranges_off_len, n_ranges};
    for (int k = 0; k < n_ranges; k++)
        if ((pi.ranges[k].off & 3u) || (pi.ranges[k].len & 3u) || (uint64_t)pi.ranges[k].off + pi.ranges[k].len > bytes)
            return fail(e, JG_ERR_INVALID, "jg_layout_patch: range out of bounds / not a multiple of 4");
    return upload_common(e, image, bytes, epoch, false, nullptr, &pi);
}
uint64_t jg_engine_layout_epoch(const JgEngine* e) { return e ? e->epoch : 0; }

int jg_item_upload(JgEngine* e, int item_id, const float* verts_xy, int n_verts, const float* poi_xy,
                   const float* poles_xyr, int n_poles, const float* piers, int n_piers) {
    if (!e || !verts_xy || n_verts < 3 || !poi_xy || n_poles < 0 || n_piers < 0) return JG_ERR_INVALID;
    if (n_verts > 4096) return fail(e, JG_ERR_CAPACITY, "item has more than 4096 vertices");
    JG_CUDA(e, cudaSetDevice(e->device));
    int rc = jg_sync(e);
    if (rc != JG_OK) return rc;
    ItemDev& it = e->items[item_id];
    cudaFree(it.verts);
    cudaFree(it.poles);
    cudaFree(it.piers);
    cudaFree(it.rots);
    cudaFree(it.rot_ext);
    it = ItemDev{};
    it.n_verts = n_verts;
    it.h_verts.assign(verts_xy, verts_xy + 2 * (size_t)n_verts);
    it.poi_x = poi_xy[0];
    it.poi_y = poi_xy[1];
    it.n_poles = n_poles;
    it.n_piers = n_piers;
    JG_CUDA(e, cudaMalloc(&it.verts, (size_t)n_verts * sizeof(float2)));
    JG_CUDA(e, cudaMemcpy(it.verts, verts_xy, (size_t)n_verts * sizeof(float2), cudaMemcpyHostToDevice));
    std::vector<float4> tmp;
    if (n_poles) {
        tmp.resize(n_poles);
        for (int i = 0; i < n_poles; i++) tmp[i] = make_float4(poles_xyr[3 * i], poles_xyr[3 * i + 1], poles_xyr[3 * i + 2], 0.f);
        JG_CUDA(e, cudaMalloc(&it.poles, (size_t)n_poles * sizeof(float4)));
        JG_CUDA(e, cudaMemcpy(it.poles, tmp.data(), (size_t)n_poles * sizeof(float4), cudaMemcpyHostToDevice));
    }
    if (n_piers) {
        JG_CUDA(e, cudaMalloc(&it.piers, (size_t)n_piers * sizeof(float4)));
        JG_CUDA(e, cudaMemcpy(it.piers, piers, (size_t)n_piers * sizeof(float4), cudaMemcpyHostToDevice));
    }
    return JG_OK;
}

int jg_item_set_rotations(JgEngine* e, int item_id, const float* rot_cos_sin, int n_rot) {
    if (!e || !rot_cos_sin || n_rot < 1 || n_rot > 256) return JG_ERR_INVALID;
    auto f = e->items.find(item_id);
    if (f == e->items.end()) return fail(e, JG_ERR_INVALID, "unknown item id");
    JG_CUDA(e, cudaSetDevice(e->device));
    int rc = jg_sync(e);
    if (rc != JG_OK) return rc;
    ItemDev& it = f->second;
    cudaFree(it.rots);
    it.rots = nullptr;
    it.n_rots = 0;
    JG_CUDA(e, cudaMalloc(&it.rots, (size_t)n_rot * sizeof(float2)));
    JG_CUDA(e, cudaMemcpy(it.rots, rot_cos_sin, (size_t)n_rot * sizeof(float2), cudaMemcpyHostToDevice));
    it.n_rots = n_rot;
    // extents of the rotated shape, with xf_point's own arithmetic (two products, one sum, no contraction): the bbox
    // of the placed shape is (ext + translation) by monotonicity of the final rounding
    std::vector<float2> ext((size_t)n_rot);
    for (int r = 0; r < n_rot; r++) {
        const Xf xf{rot_cos_sin[2 * r], rot_cos_sin[2 * r + 1], 0.0f, 0.0f};
        const float m01 = -xf.s;
        float xm = -3.402823466e+38f, ym = -3.402823466e+38f;
        for (int i = 0; i < it.n_verts; i++) {
            const float x = it.h_verts[2 * i], y = it.h_verts[2 * i + 1];
            xm = fmaxf(xm, xf.c * x + m01 * y);
            ym = fmaxf(ym, xf.s * x + xf.c * y);
        }
        ext[r] = make_float2(xm, ym);
    }
    cudaFree(it.rot_ext);
    it.rot_ext = nullptr;
    JG_CUDA(e, cudaMalloc(&it.rot_ext, (size_t)n_rot * sizeof(float2)));
    JG_CUDA(e, cudaMemcpy(it.rot_ext, ext.data(), (size_t)n_rot * sizeof(float2), cudaMemcpyHostToDevice));
    return JG_OK;
}

int jg_filter_upload(JgEngine* e, int filter_id, const uint32_t* bits, int words) {
    if (!e || filter_id < 0 || !bits || words <= 0) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    int rc = jg_sync(e);
    if (rc != JG_OK) return rc;
    FilterDev& f = e->filters[filter_id];
    cudaFree(f.bits);
    cudaFree(f.kinds);
    f = FilterDev{};
    JG_CUDA(e, cudaMalloc(&f.bits, (size_t)words * 4));
    JG_CUDA(e, cudaMemcpy(f.bits, bits, (size_t)words * 4, cudaMemcpyHostToDevice));
    f.words = words;
    return JG_OK;
}

int jg_filter_upload_keys(JgEngine* e, int filter_id, const uint64_t* hkeys, int n_keys) {
    if (!e || filter_id < 0 || n_keys < 0 || (n_keys > 0 && !hkeys)) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    int rc = jg_sync(e);
    if (rc != JG_OK) return rc;
    FilterDev& f = e->filters[filter_id];
    cudaFree(f.bits);
    cudaFree(f.kinds);
    f = FilterDev{};
    f.keyed = true;
    f.keys.assign(hkeys, hkeys + n_keys);
    return JG_OK;  // the bitset is derived against the layout in find_filter
}

// ---- queries ------------------------------------------------------------------------------------------
static int check_ready(JgEngine* e) {
    if (!e) return JG_ERR_INVALID;
    if (!e->has_layout) return fail(e, JG_ERR_NO_LAYOUT, "no layout uploaded");
    return JG_OK;
}
static int find_item(JgEngine* e, int item_id, const ItemDev** it) {
    auto f = e->items.find(item_id);
    if (f == e->items.end()) return fail(e, JG_ERR_INVALID, "unknown item id");
    *it = &f->second;
    return JG_OK;
}
static int find_filter(JgEngine* e, int filter_id, const uint32_t** d_filt, const uint8_t** d_kinds, int* grid_ok) {
    *d_filt = nullptr;
    *d_kinds = e->d_kinds;
    *grid_ok = 1;
    if (filter_id < 0) return JG_OK;
    auto g = e->filters.find(filter_id);
    if (g == e->filters.end()) return fail(e, JG_ERR_INVALID, "unknown filter id");
    FilterDev& f = g->second;
    if (f.keyed && f.bits_version != e->layout_version) {
        // HazKeyFilter semantics (filter.rs:12-58): a key hides its slot only while the slot still carries the key's version
        JG_CUDA(e, cudaSetDevice(e->device));
        if (!e->slot_versions_valid) {  // the layout came as a device blob: fetch the slot versions once
            JG_CUDA(e, cudaStreamSynchronize(e->st_up));
            e->slot_versions.resize(e->hdr.n_slots);
            if (e->hdr.n_slots)
                JG_CUDA(e, cudaMemcpy(e->slot_versions.data(), (const char*)e->d_blob + e->hdr.off_haz_version, (size_t)e->hdr.n_slots * 4,
                                      cudaMemcpyDeviceToHost));
            e->slot_versions_valid = true;
        }
        const int words = (int)((e->hdr.n_slots + 31) / 32) > 0 ? (int)((e->hdr.n_slots + 31) / 32) : 1;
        std::vector<uint32_t> hb((size_t)words, 0u);
        for (uint64_t k : f.keys) {
            const uint32_t slot = (uint32_t)k, ver = (uint32_t)(k >> 32);
            if (slot < e->hdr.n_slots && e->slot_versions[slot] == ver) hb[slot >> 5] |= 1u << (slot & 31);
        }
        JG_CUDA(e, cudaStreamSynchronize(e->compute));  // batches in flight may still read the old bitset
        for (ChunkCtx& c : e->chunks) JG_CUDA(e, cudaStreamSynchronize(c.st));
        if (f.words < words) {
            cudaFree(f.bits);
            f.bits = nullptr;
            f.words = 0;
            JG_CUDA(e, cudaMalloc(&f.bits, (size_t)words * 4 * 2));
            f.words = words * 2;
        }
        JG_CUDA(e, cudaMemset(f.bits, 0, (size_t)f.words * 4));
        JG_CUDA(e, cudaMemcpy(f.bits, hb.data(), (size_t)words * 4, cudaMemcpyHostToDevice));
        f.bits_version = e->layout_version;
        f.kinds_version = 0;  // the node summary depends on the bitset
    }
    if ((size_t)f.words * 32 < e->hdr.n_slots)
        return fail(e, JG_ERR_CAPACITY, "filter bitset smaller than the hazard slot capacity");
    if (f.kinds_version != e->layout_version) {  // node summary of this filter for the current layout
        JG_CUDA(e, cudaSetDevice(e->device));
        const size_t need = detect_kinds_bytes((int)e->hdr.n_nodes);
        if (f.kinds_cap < need) {
            cudaFree(f.kinds);
            f.kinds = nullptr;
            f.kinds_cap = 0;
            JG_CUDA(e, cudaMalloc(&f.kinds, need * 2));
            f.kinds_cap = need * 2;
        }
        JG_CUDA(e, launch_node_kinds(e->L, (int)e->hdr.root_level, f.bits, f.kinds, e->compute));
        // Under a filter, nodes whose Entire entries are all hidden become leaves of the summary: k_grid_build may refuse
        // what the NoFilter summary accepted (leaf record limits). Its verdict decides this filter's kernel path.
        JG_CUDA(e, detect_grid_usable(f.kinds, (int)e->hdr.n_nodes, e->compute, &f.grid_ok));
        e->stat_launches += 2;
        f.kinds_version = e->layout_version;
    }
    *d_filt = f.bits;
    *d_kinds = f.kinds;
    *grid_ok = f.grid_ok;
    return JG_OK;
}

// A detect batch is a sequence of segments (item_ids[k] for queries seg_offsets[k] .. seg_offsets[k+1]).
// It is cut into launch pieces of at most `max_q` queries and JG_MAX_SEG segments; each piece gets a
// SegTable with offsets relative to the piece.
struct Piece {
    int64_t off, m;
    SegTable tab;
};
struct PieceCursor {
    int64_t q = 0;
    int seg = 0;
};
static bool next_piece(const JgEngine* e, int n_seg, const ItemDev* const* items, const int64_t* seg_offsets, int64_t max_q,
                       PieceCursor& cur, Piece& p) {
    const int64_t total = seg_offsets[n_seg];
    while (cur.seg < n_seg && seg_offsets[cur.seg + 1] <= cur.q) cur.seg++;
    if (cur.q >= total || cur.seg >= n_seg) return false;
    p.off = cur.q;
    p.tab.n_seg = 0;
    int64_t covered = 0, edge_pairs = 0;
    int sgi = cur.seg;
    while (sgi < n_seg && p.tab.n_seg < JG_MAX_SEG && covered < max_q) {
        const int64_t b = seg_offsets[sgi] > cur.q ? seg_offsets[sgi] : cur.q;
        int64_t len = seg_offsets[sgi + 1] - b;
        if (len <= 0) { sgi++; continue; }
        const ItemDev& it = *items[sgi];
        if (len > max_q - covered) len = max_q - covered;
        // the flat (survivor, edge) index of k_s2_edges must fit an int
        const int64_t room = (0x7FFFFFFFLL - edge_pairs) / it.n_verts;
        if (len > room) len = room;
        if (len <= 0) break;
        SegDesc& sd = p.tab.seg[p.tab.n_seg++];
        sd.begin = (int)covered;
        sd.n_verts = it.n_verts;
        sd.n_poles = it.n_poles < e->cfg.n_ff_poles ? it.n_poles : e->cfg.n_ff_poles;
        sd.n_piers = it.n_piers;
        sd.verts = it.verts;
        sd.poles = it.poles;
        sd.piers = it.piers;
        sd.poi_x = it.poi_x;
        sd.poi_y = it.poi_y;
        sd.rots = it.rots;
        sd.n_rots = it.n_rots;
        sd.gbeg = 0;
        covered += len;
        edge_pairs += len * it.n_verts;
        if (b + len < seg_offsets[sgi + 1]) break;  // segment cut: the rest goes to the next piece
        sgi++;
    }
    p.m = covered;
    p.tab.n_queries = (int)covered;
    cur.q += covered;
    return covered > 0;
}

static int resolve_segments(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets,
                            std::vector<const ItemDev*>& items) {
    if (n_seg <= 0 || !item_ids || !seg_offsets || seg_offsets[0] != 0) return fail(e, JG_ERR_INVALID, "bad segment table");
    items.resize(n_seg);
    for (int k = 0; k < n_seg; k++) {
        if (seg_offsets[k + 1] < seg_offsets[k]) return fail(e, JG_ERR_INVALID, "segment offsets must be non-decreasing");
        int rc = find_item(e, item_ids[k], &items[k]);
        if (rc != JG_OK) return rc;
    }
    return JG_OK;
}

#define JG_DEVICE_PIECE_MIN (1 << 16)

// one piece on one stream: the fused kernel when the grid serves it, else the two-kernel sequence
static cudaError_t launch_piece(JgEngine* e, const DetectPlan& plan, const Piece& p, const XfSrc& src, const uint32_t* d_filt, const uint8_t* d_kinds,
                                int mode, uint8_t* d_out, const DetectScratch& sc, cudaStream_t st, int* launches) {
    if (e->use_fused && detect_fused_applicable(plan, p.tab, mode)) {
        cudaError_t err = cudaMemsetAsync(sc.ctl, 0, sizeof(DetectCtl), st);
        if (err != cudaSuccess) return err;
        return launch_detect_fused(e->L, p.tab, src, d_filt, d_kinds, d_out, sc, e->n_sms, 1, st, launches);
    }
    return launch_detect(e->L, plan, p.tab, src, d_filt, d_kinds, mode, d_out, sc, e->n_sms, st, launches);
}
// the launch plan of a batch: the engine's, minus the grid kernels when this filter's summary does not admit them
static DetectPlan plan_for(const JgEngine* e, int filter_grid_ok) {
    DetectPlan plan = e->plan;
    if (!filter_grid_ok) plan.use_grid = 0;
    return plan;
}
static int detect_multi_device(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets,
                               const float* d_transforms, int filter_id, int mode, uint8_t* d_out) {
    int rc = check_ready(e);
    if (rc != JG_OK) return rc;
    std::vector<const ItemDev*> items;
    rc = resolve_segments(e, n_seg, item_ids, seg_offsets, items);
    if (rc != JG_OK) return rc;
    const uint32_t* d_filt;
    const uint8_t* d_kinds;
    int fgrid = 1;
    rc = find_filter(e, filter_id, &d_filt, &d_kinds, &fgrid);
    if (rc != JG_OK) return rc;
    const DetectPlan plan = plan_for(e, fgrid);
    const int64_t n = seg_offsets[n_seg];
    if (n > 0 && (!d_transforms || !d_out)) return JG_ERR_INVALID;
    if (n == 0) return JG_OK;
    JG_CUDA(e, cudaSetDevice(e->device));
    int launches = 0;
    JG_CUDA(e, cudaEventRecord(e->ev0, e->compute));
    PieceCursor cur;
    Piece p;
    static const int n_pieces = [] { const char* v = getenv("JG_DEVICE_PIECES"); int k = v ? atoi(v) : 1; return k < 1 ? 1 : k; }();
    if (n_pieces > 1 && n >= 2 * JG_DEVICE_PIECE_MIN && e->chunks.size() >= 2) {
        // Large device-resident batch: a few pieces on the chunk streams, so that the tail of one piece's
        // persistent kernels (SMs draining their last work items) is filled by the next piece's kernels.
        int64_t piece = (n + n_pieces - 1) / n_pieces;
        if (piece < JG_DEVICE_PIECE_MIN) piece = JG_DEVICE_PIECE_MIN;
        if (piece > JG_MAX_LAUNCH_QUERIES) piece = JG_MAX_LAUNCH_QUERIES;
        for (ChunkCtx& c : e->chunks) {
            if (!c.done) JG_CUDA(e, cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
            JG_CUDA(e, cudaStreamWaitEvent(c.st, e->ev0, 0));
        }
        size_t k = 0;
        while (next_piece(e, n_seg, items.data(), seg_offsets, piece, cur, p)) {
            ChunkCtx& c = e->chunks[k++ % e->chunks.size()];
            rc = ensure_ctx(e, c, (size_t)piece, 0, false);
            if (rc != JG_OK) return rc;
            JG_CUDA(e, launch_piece(e, plan, p, XfSrc{(const float4*)d_transforms + p.off, nullptr, nullptr}, d_filt, d_kinds, mode, d_out + p.off, c.sc,
                                    c.st, &launches));
        }
        for (ChunkCtx& c : e->chunks) {
            JG_CUDA(e, cudaEventRecord(c.done, c.st));
            JG_CUDA(e, cudaStreamWaitEvent(e->compute, c.done, 0));
        }
    } else {
        const int64_t max_q = JG_MAX_LAUNCH_QUERIES;
        rc = ensure_ctx(e, e->main_ctx, (size_t)(n < max_q ? n : max_q), 0, false);
        if (rc != JG_OK) return rc;
        while (next_piece(e, n_seg, items.data(), seg_offsets, max_q, cur, p))
            JG_CUDA(e, launch_piece(e, plan, p, XfSrc{(const float4*)d_transforms + p.off, nullptr, nullptr}, d_filt, d_kinds, mode, d_out + p.off,
                                    e->main_ctx.sc, e->compute, &launches));
    }
    JG_CUDA(e, cudaEventRecord(e->ev1, e->compute));
    e->timing_valid = true;
    e->stat_launches += launches;
    e->stat_queries += (uint64_t)n;
    return JG_OK;
}

// Host-pointer batches: pieces of at most chunk_queries queries, piece k on stream k mod n_streams, so
// the H2D copy of piece k+1 overlaps the kernels of piece k and the D2H copy of piece k-1.
// `transforms` (16 B/query) or, when it is null, the compact form `rot_idx` + `txy` (9 B/query)
static int detect_multi_host(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                             const uint8_t* rot_idx, const float* txy, int filter_id, int mode, uint8_t* out, bool wait,
                             const JgLattice* lat = nullptr, const uint16_t* ij = nullptr) {
    int rc = check_ready(e);
    if (rc != JG_OK) return rc;
    std::vector<const ItemDev*> items;
    rc = resolve_segments(e, n_seg, item_ids, seg_offsets, items);
    if (rc != JG_OK) return rc;
    const uint32_t* d_filt;
    const uint8_t* d_kinds;
    int fgrid = 1;
    rc = find_filter(e, filter_id, &d_filt, &d_kinds, &fgrid);
    if (rc != JG_OK) return rc;
    const DetectPlan plan = plan_for(e, fgrid);
    const int64_t n = seg_offsets[n_seg];
    const bool indexed = transforms == nullptr;
    const bool lattice = indexed && ij != nullptr;
    if (n > 0 && (!out || (indexed && (!rot_idx || (!txy && !lattice))) || (lattice && !lat))) return JG_ERR_INVALID;
    if (indexed)
        for (int k = 0; k < n_seg; k++)
            if (seg_offsets[k + 1] > seg_offsets[k] && items[k]->n_rots == 0)
                return fail(e, JG_ERR_INVALID, "indexed queries need jg_item_set_rotations for every item of the batch");
    if (n == 0) return JG_OK;
    JG_CUDA(e, cudaSetDevice(e->device));
    if (e->use_fused && !lattice && n >= JG_STREAM_MIN_QUERIES && n <= JG_MAX_LAUNCH_QUERIES && n_seg <= JG_MAX_SEG) {
        // Whole batch as ONE fused kernel that follows the copy front (see k_detect_fused).
        PieceCursor cur0;
        Piece whole;
        if (next_piece(e, n_seg, items.data(), seg_offsets, n, cur0, whole) && whole.m == n &&
            detect_fused_applicable(plan, whole.tab, mode)) {
            ChunkCtx& c = e->stream_ctx;
            JG_CUDA(e, cudaStreamSynchronize(e->compute));  // the previous streamed batch is done with the buffers
            JG_CUDA(e, cudaStreamSynchronize(e->st_copy));
            rc = ensure_ctx(e, c, (size_t)n, (size_t)n, true);
            if (rc != JG_OK) return rc;
            int64_t sc_q = (n + JG_STREAM_MAX_CHUNKS - 1) / JG_STREAM_MAX_CHUNKS;
            if (sc_q < JG_STREAM_CHUNK) sc_q = JG_STREAM_CHUNK;
            sc_q = (sc_q + 31) & ~(int64_t)31;
            float2* d_txy = (float2*)c.d_xf;
            uint8_t* d_ridx = (uint8_t*)(d_txy + n);
            const XfSrc src = indexed ? XfSrc{nullptr, d_ridx, d_txy} : XfSrc{c.d_xf, nullptr, nullptr};
            JG_CUDA(e, cudaMemsetAsync(c.sc.ctl, 0, sizeof(DetectCtl), e->compute));
            JG_CUDA(e, cudaEventRecord(e->ev_reset, e->compute));
            int launches = 0;
            JG_CUDA(e, launch_detect_fused(e->L, whole.tab, src, d_filt, d_kinds, c.d_out, c.sc, e->n_sms, 0, e->compute, &launches));
            JG_CUDA(e, cudaMemcpyAsync(out, c.d_out, (size_t)n, cudaMemcpyDeviceToHost, e->compute));
            JG_CUDA(e, cudaMemcpyAsync(e->h_err, &c.sc.ctl->error, sizeof(int), cudaMemcpyDeviceToHost, e->compute));
            JG_CUDA(e, cudaStreamWaitEvent(e->st_copy, e->ev_reset, 0));  // avail is bumped only after the reset
            int k = 0;
            for (int64_t off = 0; off < n; off += sc_q, k++) {
                const int64_t m = (n - off) < sc_q ? (n - off) : sc_q;
                if (!indexed) {
                    JG_CUDA(e, cudaMemcpyAsync(c.d_xf + off, transforms + 4 * off, (size_t)m * sizeof(float4), cudaMemcpyHostToDevice, e->st_copy));
                } else {
                    JG_CUDA(e, cudaMemcpyAsync(d_txy + off, txy + 2 * off, (size_t)m * sizeof(float2), cudaMemcpyHostToDevice, e->st_copy));
                    JG_CUDA(e, cudaMemcpyAsync(d_ridx + off, rot_idx + off, (size_t)m, cudaMemcpyHostToDevice, e->st_copy));
                }
                e->h_avail[k] = (int)(off + m);
                JG_CUDA(e, cudaMemcpyAsync(&c.sc.ctl->avail, &e->h_avail[k], sizeof(int), cudaMemcpyHostToDevice, e->st_copy));
            }
            e->stream_pending = true;
            e->stat_launches += launches;
            e->stat_h2d += (uint64_t)n * (indexed ? 9 : sizeof(float4));
            e->stat_d2h += (uint64_t)n;
            e->stat_queries += (uint64_t)n;
            if (wait) return jg_sync(e);
            return JG_OK;
        }
    }
    // Piece schedule. The persistent kernels of consecutive pieces cannot overlap each other (each fills every SM slot) and a
    // launch pair has ~45 us of fixed latency (tools/size_sweep.py: 64 Ki queries 53 us, 256 Ki 87, 1 Mi 200), so:
    //  * a batch of up to two chunks (512 Ki queries) goes as ONE piece (262 Ki queries: 46 us of copy + 93 us of kernels in
    //    one piece, 23 + 70 + 70 us in two);
    //  * engine idle (nothing in flight on the piece streams - the caller runs one batch at a time): pieces of one chunk.
    //    With 9 B/query the copy (0.16 us per Ki queries) and the kernels (0.155) run at the same rate, so the step costs
    //    about copy(first piece) + the kernel chain; 4 x 256 Ki = 0.384 ms per 1 Mi queries. Measured and not better
    //    (tools/schedule_probe.py, tools/trace_e2e.py, profiles/r2u_*): a short first piece and the rest in one
    //    (256 Ki + 768 Ki: 0.417 ms - all copies must land before the last piece's 173 us of kernels start), one piece
    //    (0.43 ms), 512 Ki pieces;
    //  * batches already in flight (the caller pipelines): the copies of this batch hide behind the kernels of the previous
    //    ones, so whole pieces of up to four chunks - one launch pair per 1 Mi queries: 0.195 ms per 1 Mi batch instead of
    //    0.259 ms in 256 Ki pieces.
    // jg_engine_set_batching(chunk > 0) pins fixed pieces of that size in both situations.
    std::vector<int64_t> sched;
    {
        const int64_t cq = e->chunk_queries;
        int64_t piece = cq;
        if (e->auto_schedule) {
            bool busy = false;
            for (ChunkCtx& c : e->chunks) busy = busy || cudaStreamQuery(c.st) == cudaErrorNotReady;
            (void)cudaGetLastError();
            if (busy) piece = 4 * cq;
        }
        if (n <= 2 * cq) piece = n;
        for (int64_t done = 0; done < n; done += piece) sched.push_back(piece < n - done ? piece : n - done);
    }
    int64_t max_piece = 0;
    for (int64_t m : sched) max_piece = m > max_piece ? m : max_piece;
    // (the slots are sized for the larger of the two schedules: the same batch size must not re-allocate - a device-wide
    // synchronisation - when the caller goes from one batch at a time to batches in flight)
    if (e->auto_schedule && max_piece < (n < 4 * e->chunk_queries ? n : 4 * e->chunk_queries)) max_piece = n < 4 * e->chunk_queries ? n : 4 * e->chunk_queries;
    size_t sched_k = 0;
    PieceCursor cur;
    Piece p;
    // (a piece can come out shorter than asked - segment table full -: what is left then goes in pieces of max_piece)
    auto piece_size = [&](int64_t) { return sched_k < sched.size() ? sched[sched_k++] : max_piece; };
    const bool tr = trace_on();
    if (tr) {
        e->trace_used = 0;
        e->trace_m.clear();
        trace_mark(e, e->chunks[e->next_chunk % e->chunks.size()].st);
    }
    // size every piece slot now (no-ops once they are): a slot that first grows in the middle of a run of batches would
    // stall that batch by several milliseconds of cudaMalloc
    for (ChunkCtx& c : e->chunks)
        if (c.cap_q < (size_t)max_piece || c.cap_sv < (size_t)max_piece || c.cap_out < (size_t)max_piece) {
            JG_CUDA(e, cudaStreamSynchronize(c.st));
            rc = ensure_ctx(e, c, (size_t)max_piece, (size_t)max_piece, true);
            if (rc != JG_OK) return rc;
        }
    while (next_piece(e, n_seg, items.data(), seg_offsets, piece_size(cur.q), cur, p)) {
        ChunkCtx& c = e->chunks[e->next_chunk++ % e->chunks.size()];
        JG_CUDA(e, cudaStreamSynchronize(c.st));  // buffers of this slot are free again
        if (tr) trace_mark(e, c.st), e->trace_m.push_back(p.m);
        rc = ensure_ctx(e, c, (size_t)max_piece, (size_t)max_piece, true);
        if (rc != JG_OK) return rc;
        XfSrc src{c.d_xf, nullptr, nullptr};
        if (!indexed) {
            JG_CUDA(e, cudaMemcpyAsync(c.d_xf, transforms + 4 * p.off, (size_t)p.m * sizeof(float4), cudaMemcpyHostToDevice, c.st));
            e->stat_h2d += (uint64_t)p.m * sizeof(float4);
        } else if (lattice) {  // lattice coordinates (2 x u16) followed by the rotation indices
            ushort2* d_ij = (ushort2*)c.d_xf;
            uint8_t* d_ridx = (uint8_t*)(d_ij + p.m);
            JG_CUDA(e, cudaMemcpyAsync(d_ij, ij + 2 * p.off, (size_t)p.m * sizeof(ushort2), cudaMemcpyHostToDevice, c.st));
            JG_CUDA(e, cudaMemcpyAsync(d_ridx, rot_idx + p.off, (size_t)p.m, cudaMemcpyHostToDevice, c.st));
            src = XfSrc{nullptr, d_ridx, nullptr, d_ij, make_float4(lat->x0, lat->y0, lat->dx, lat->dy)};
            e->stat_h2d += (uint64_t)p.m * 5;
        } else {  // the staging buffer holds the translations followed by the rotation indices
            float2* d_txy = (float2*)c.d_xf;
            uint8_t* d_ridx = (uint8_t*)(d_txy + p.m);
            JG_CUDA(e, cudaMemcpyAsync(d_txy, txy + 2 * p.off, (size_t)p.m * sizeof(float2), cudaMemcpyHostToDevice, c.st));
            JG_CUDA(e, cudaMemcpyAsync(d_ridx, rot_idx + p.off, (size_t)p.m, cudaMemcpyHostToDevice, c.st));
            src = XfSrc{nullptr, d_ridx, d_txy};
            e->stat_h2d += (uint64_t)p.m * 9;
        }
        int launches = 0;
        if (tr) trace_mark(e, c.st);
        JG_CUDA(e, launch_piece(e, plan, p, src, d_filt, d_kinds, mode, c.d_out, c.sc, c.st, &launches));
        if (tr) trace_mark(e, c.st);
        JG_CUDA(e, cudaMemcpyAsync(out + p.off, c.d_out, (size_t)p.m, cudaMemcpyDeviceToHost, c.st));
        if (tr) trace_mark(e, c.st);
        e->stat_launches += launches;
        e->stat_d2h += (uint64_t)p.m;
    }
    if (wait)
        for (ChunkCtx& c : e->chunks) JG_CUDA(e, cudaStreamSynchronize(c.st));
    e->stat_queries += (uint64_t)n;
    return JG_OK;
}

int jg_detect_multi(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                    int filter_id, int mode, uint8_t* out) {
    return detect_multi_host(e, n_seg, item_ids, seg_offsets, transforms, nullptr, nullptr, filter_id, mode, out, true);
}
int jg_detect_multi_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                          int filter_id, int mode, uint8_t* out) {
    return detect_multi_host(e, n_seg, item_ids, seg_offsets, transforms, nullptr, nullptr, filter_id, mode, out, false);
}
int jg_detect_multi_indexed(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                            const float* txy, int filter_id, int mode, uint8_t* out) {
    if (!rot_idx || !txy) return JG_ERR_INVALID;
    return detect_multi_host(e, n_seg, item_ids, seg_offsets, nullptr, rot_idx, txy, filter_id, mode, out, true);
}
int jg_detect_multi_indexed_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                                  const float* txy, int filter_id, int mode, uint8_t* out) {
    if (!rot_idx || !txy) return JG_ERR_INVALID;
    return detect_multi_host(e, n_seg, item_ids, seg_offsets, nullptr, rot_idx, txy, filter_id, mode, out, false);
}
int jg_detect_multi_lattice(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                            const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out) {
    if (!lattice || !rot_idx || !ij) return JG_ERR_INVALID;
    return detect_multi_host(e, n_seg, item_ids, seg_offsets, nullptr, rot_idx, nullptr, filter_id, mode, out, true, lattice, ij);
}
int jg_detect_multi_lattice_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                                  const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out) {
    if (!lattice || !rot_idx || !ij) return JG_ERR_INVALID;
    return detect_multi_host(e, n_seg, item_ids, seg_offsets, nullptr, rot_idx, nullptr, filter_id, mode, out, false, lattice, ij);
}
void jg_lattice_points(const JgLattice* lattice, const uint16_t* ij, int64_t n, float* txy) {
    if (!lattice || !ij || !txy) return;
    for (int64_t q = 0; q < n; q++) {
        txy[2 * q] = lattice_coord(lattice->x0, lattice->dx, ij[2 * q]);
        txy[2 * q + 1] = lattice_coord(lattice->y0, lattice->dy, ij[2 * q + 1]);
    }
}
int jg_detect_multi_device(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* d_transforms,
                           int filter_id, int mode, uint8_t* d_out) {
    return detect_multi_device(e, n_seg, item_ids, seg_offsets, d_transforms, filter_id, mode, d_out);
}
int jg_detect_batch(JgEngine* e, int item_id, const float* transforms, int64_t n, int filter_id, int mode, uint8_t* out) {
    if (n < 0) return JG_ERR_INVALID;
    const int64_t offs[2] = {0, n};
    return detect_multi_host(e, 1, &item_id, offs, transforms, nullptr, nullptr, filter_id, mode, out, true);
}
int jg_detect_batch_async(JgEngine* e, int item_id, const float* transforms, int64_t n, int filter_id, int mode,
                          uint8_t* out) {
    if (n < 0) return JG_ERR_INVALID;
    const int64_t offs[2] = {0, n};
    return detect_multi_host(e, 1, &item_id, offs, transforms, nullptr, nullptr, filter_id, mode, out, false);
}
int jg_detect_batch_device(JgEngine* e, int item_id, const float* d_transforms, int64_t n, int filter_id, int mode,
                           uint8_t* d_out) {
    if (n < 0) return JG_ERR_INVALID;
    const int64_t offs[2] = {0, n};
    return detect_multi_device(e, 1, &item_id, offs, d_transforms, filter_id, mode, d_out);
}

static int collect_checks(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const void* t, const void* out, int words,
                          std::vector<const ItemDev*>& items) {
    int rc = check_ready(e);
    if (rc != JG_OK) return rc;
    rc = resolve_segments(e, n_seg, item_ids, seg_offsets, items);
    if (rc != JG_OK) return rc;
    if (words <= 0 || (seg_offsets[n_seg] > 0 && (!t || !out))) return JG_ERR_INVALID;
    if ((size_t)words * 32 < e->hdr.n_slots) return fail(e, JG_ERR_CAPACITY, "collector words do not cover the hazard slot capacity");
    if (words > 1024) return fail(e, JG_ERR_CAPACITY, "collector wider than 1024 words");
    return JG_OK;
}

// one launch piece of a collect batch: ONE k_collect_grid launch for all of its segments where the grid serves it, else the
// register-walk kernel segment by segment
static cudaError_t launch_collect_piece(JgEngine* e, Piece& p, const float4* d_xf, int mode, const uint32_t* d_seed, uint32_t* d_out, int words,
                                        int* d_next, cudaStream_t st, int* launches) {
    if (!e->force_walk && collect_grid_applicable(e->plan, p.tab, words))
        return launch_collect_grid(e->L, p.tab, e->d_kinds, d_xf, mode, d_seed, d_out, words, d_next, e->n_sms, st, launches);
    for (int k = 0; k < p.tab.n_seg; k++) {
        const SegDesc& sd = p.tab.seg[k];
        const int end = k + 1 < p.tab.n_seg ? p.tab.seg[k + 1].begin : p.tab.n_queries;
        DevItem di;
        di.verts = sd.verts, di.n_verts = sd.n_verts, di.poi_x = sd.poi_x, di.poi_y = sd.poi_y;
        di.poles = sd.poles, di.n_poles = sd.n_poles, di.piers = sd.piers, di.n_piers = sd.n_piers;
        const int group = e->collect_group ? e->collect_group : auto_group(sd.n_verts);
        cudaError_t err = launch_collect(e->L, di, d_xf + sd.begin, end - sd.begin, mode, d_seed ? d_seed + (size_t)sd.begin * words : nullptr,
                                         d_out + (size_t)sd.begin * words, words, group, e->n_sms, st, launches);
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

static int collect_multi_device(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* d_transforms, int mode,
                                const uint32_t* d_seed, uint32_t* d_out_bits, int words) {
    std::vector<const ItemDev*> items;
    int rc = collect_checks(e, n_seg, item_ids, seg_offsets, d_transforms, d_out_bits, words, items);
    if (rc != JG_OK) return rc;
    const int64_t n = seg_offsets[n_seg];
    if (n == 0) return JG_OK;
    JG_CUDA(e, cudaSetDevice(e->device));
    int launches = 0;
    JG_CUDA(e, cudaEventRecord(e->ev0, e->compute));
    rc = ensure_ctx(e, e->main_ctx, 1, 0, false);  // the work counter lives in the scratch block
    if (rc != JG_OK) return rc;
    PieceCursor cur;
    Piece p;
    while (next_piece(e, n_seg, items.data(), seg_offsets, JG_MAX_LAUNCH_QUERIES, cur, p))
        JG_CUDA(e, launch_collect_piece(e, p, (const float4*)d_transforms + p.off, mode, d_seed ? d_seed + (size_t)p.off * words : nullptr,
                                        d_out_bits + (size_t)p.off * words, words, &e->main_ctx.sc.ctl->s1_next, e->compute, &launches));
    JG_CUDA(e, cudaEventRecord(e->ev1, e->compute));
    e->timing_valid = true;
    e->stat_launches += launches;
    e->stat_queries += (uint64_t)n;
    return JG_OK;
}

static int collect_multi_host(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                              const uint32_t* seed_bits, uint32_t* out_bits, int words, bool wait) {
    std::vector<const ItemDev*> items;
    int rc = collect_checks(e, n_seg, item_ids, seg_offsets, transforms, out_bits, words, items);
    if (rc != JG_OK) return rc;
    const int64_t n = seg_offsets[n_seg];
    if (n == 0) return JG_OK;
    JG_CUDA(e, cudaSetDevice(e->device));
    const size_t out_per_q = (size_t)words * 4;
    const int64_t chunk = e->chunk_queries;
    const size_t slot_q = (size_t)(chunk < n ? chunk : n);
    PieceCursor cur;
    Piece p;
    while (next_piece(e, n_seg, items.data(), seg_offsets, chunk, cur, p)) {
        ChunkCtx& c = e->chunks[e->next_chunk++ % e->chunks.size()];
        JG_CUDA(e, cudaStreamSynchronize(c.st));
        rc = ensure_ctx(e, c, slot_q, slot_q * out_per_q, true);
        if (rc != JG_OK) return rc;
        JG_CUDA(e, cudaMemcpyAsync(c.d_xf, transforms + 4 * p.off, (size_t)p.m * sizeof(float4), cudaMemcpyHostToDevice, c.st));
        const uint32_t* d_seed = nullptr;
        if (seed_bits) {  // seeded in place: a warp reads the seeds of its group before it writes that group's results
            JG_CUDA(e, cudaMemcpyAsync(c.d_out, (const char*)seed_bits + (size_t)p.off * out_per_q, (size_t)p.m * out_per_q, cudaMemcpyHostToDevice, c.st));
            d_seed = (const uint32_t*)c.d_out;
            e->stat_h2d += (uint64_t)p.m * out_per_q;
        }
        int launches = 0;
        JG_CUDA(e, launch_collect_piece(e, p, c.d_xf, mode, d_seed, (uint32_t*)c.d_out, words, &c.sc.ctl->s1_next, c.st, &launches));
        JG_CUDA(e, cudaMemcpyAsync((char*)out_bits + (size_t)p.off * out_per_q, c.d_out, (size_t)p.m * out_per_q, cudaMemcpyDeviceToHost, c.st));
        e->stat_launches += launches;
        e->stat_h2d += (uint64_t)p.m * sizeof(float4);
        e->stat_d2h += (uint64_t)p.m * out_per_q;
    }
    if (wait)
        for (ChunkCtx& c : e->chunks) JG_CUDA(e, cudaStreamSynchronize(c.st));
    e->stat_queries += (uint64_t)n;
    return JG_OK;
}

int jg_collect_multi(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                     const uint32_t* seed_bits, uint32_t* out_bits, int words) {
    return collect_multi_host(e, n_seg, item_ids, seg_offsets, transforms, mode, seed_bits, out_bits, words, true);
}
int jg_collect_multi_async(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                           const uint32_t* seed_bits, uint32_t* out_bits, int words) {
    return collect_multi_host(e, n_seg, item_ids, seg_offsets, transforms, mode, seed_bits, out_bits, words, false);
}
int jg_collect_multi_device(JgEngine* e, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* d_transforms, int mode,
                            const uint32_t* d_seed_bits, uint32_t* d_out_bits, int words) {
    return collect_multi_device(e, n_seg, item_ids, seg_offsets, d_transforms, mode, d_seed_bits, d_out_bits, words);
}
int jg_collect_batch_device(JgEngine* e, int item_id, const float* d_transforms, int64_t n, int mode, uint32_t* d_out_bits, int words) {
    if (n < 0) return JG_ERR_INVALID;
    const int64_t offs[2] = {0, n};
    return collect_multi_device(e, 1, &item_id, offs, d_transforms, mode, nullptr, d_out_bits, words);
}
int jg_collect_batch(JgEngine* e, int item_id, const float* transforms, int64_t n, int mode, uint32_t* out_bits, int words) {
    if (n < 0) return JG_ERR_INVALID;
    const int64_t offs[2] = {0, n};
    return collect_multi_host(e, 1, &item_id, offs, transforms, mode, nullptr, out_bits, words, true);
}
int jg_collect_batch_async(JgEngine* e, int item_id, const float* transforms, int64_t n, int mode, uint32_t* out_bits, int words) {
    if (n < 0) return JG_ERR_INVALID;
    const int64_t offs[2] = {0, n};
    return collect_multi_host(e, 1, &item_id, offs, transforms, mode, nullptr, out_bits, words, false);
}

int jg_engine_stats(const JgEngine* e, uint64_t out[8]) {
    if (!e || !out) return JG_ERR_INVALID;
    std::memset(out, 0, 8 * sizeof(uint64_t));
    out[0] = e->stat_launches;
    out[1] = e->stat_queries;
    out[2] = e->stat_h2d;
    out[3] = e->stat_d2h;
    out[4] = e->stat_uploads;
    out[7] = e->stat_patched;
    if (e->main_ctx.sc.ctl) {  // survivors of the last *_device detect launch sequence
        DetectCtl c;
        cudaSetDevice(e->device);
        if (cudaMemcpy(&c, e->main_ctx.sc.ctl, sizeof(c), cudaMemcpyDeviceToHost) == cudaSuccess) {
            out[5] = (uint64_t)c.n_surv;
            out[6] = (uint64_t)c.fallbacks;
        }
    }
    return JG_OK;
}

float jg_engine_last_kernel_ms(JgEngine* e) {
    if (!e || !e->timing_valid) return -1.0f;
    cudaSetDevice(e->device);
    if (cudaEventSynchronize(e->ev1) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int jg_timer_start(JgEngine* e) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JG_CUDA(e, cudaEventRecord(e->tm0, e->compute));
    return JG_OK;
}
// The stopwatch with the host taken out of the timed region: a one-thread gate kernel holds the compute stream until
// jg_timer_stop has recorded the closing event, so everything enqueued in between runs back to back on the device and the
// elapsed time is the device's alone. Without it the time the host needs to enqueue the launches counts whenever it exceeds what
// is queued ahead on the stream - on an 8-GPU box with one busy process per GPU single steps doubled (0.21 -> 0.44 ms). The spin
// is bounded (about 0.1 s), a host that never opens the gate cannot hang the device.
__global__ void k_gate(volatile int* flag) {
    const long long t0 = clock64();
    while (!*flag && clock64() - t0 < (1ll << 28)) {}
}
int jg_timer_start_gated(JgEngine* e) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    *(volatile int*)e->h_gate = 0;
    e->gate_armed = true;
    k_gate<<<1, 1, 0, e->compute>>>(e->h_gate);
    JG_CUDA(e, cudaGetLastError());
    JG_CUDA(e, cudaEventRecord(e->tm0, e->compute));
    return JG_OK;
}
int jg_timer_stop(JgEngine* e, float* ms) {
    if (!e || !ms) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JG_CUDA(e, cudaEventRecord(e->tm1, e->compute));
    if (e->gate_armed) {
        *(volatile int*)e->h_gate = 1;
        e->gate_armed = false;
    }
    JG_CUDA(e, cudaEventSynchronize(e->tm1));
    JG_CUDA(e, cudaEventElapsedTime(ms, e->tm0, e->tm1));
    return JG_OK;
}

// ---- batched placement search rounds (search.cu) -------------------------------------------------------
__global__ void k_search_fetch(unsigned long long* best, const uint8_t* ridx, const float2* txy) {
    const unsigned long long key = best[0];
    if (key == ~0ull) return;
    const uint32_t q = (uint32_t)key;
    const float2 t = txy[q];
    best[2] = ((unsigned long long)__float_as_uint(t.y) << 32) | __float_as_uint(t.x);
    best[3] = ridx[q];
}

int jg_search_round(JgEngine* e, int item_id, const JgSearchRound* round, int64_t n, int filter_id, JgSearchBest* best) {
    int rc = check_ready(e);
    if (rc != JG_OK) return rc;
    if (!round || !best || n <= 0 || n > JG_MAX_LAUNCH_QUERIES || (round->mode != 0 && round->mode != 1)) return JG_ERR_INVALID;
    const int ids[1] = {item_id};
    const int64_t offs[2] = {0, n};
    std::vector<const ItemDev*> items;
    rc = resolve_segments(e, 1, ids, offs, items);
    if (rc != JG_OK) return rc;
    const ItemDev& it = *items[0];
    if (it.n_rots == 0 || !it.rot_ext) return fail(e, JG_ERR_INVALID, "search rounds need jg_item_set_rotations (a single (1, 0) entry for items that do not rotate)");
    if (round->rot_idx >= it.n_rots || (round->mode == 1 && round->rot_idx < 0)) return fail(e, JG_ERR_INVALID, "search round: bad rotation index");
    const uint32_t* d_filt;
    const uint8_t* d_kinds;
    int fgrid = 1;
    rc = find_filter(e, filter_id, &d_filt, &d_kinds, &fgrid);
    if (rc != JG_OK) return rc;
    const DetectPlan plan = plan_for(e, fgrid);
    JG_CUDA(e, cudaSetDevice(e->device));
    if (!e->d_search) {
        JG_CUDA(e, cudaMalloc(&e->d_search, 4 * sizeof(unsigned long long)));
        JG_CUDA(e, cudaHostAlloc((void**)&e->h_search, 4 * sizeof(unsigned long long), cudaHostAllocDefault));
    }
    ChunkCtx& c = e->main_ctx;
    rc = ensure_ctx(e, c, (size_t)n, (size_t)n, true);
    if (rc != JG_OK) return rc;
    float2* d_txy = (float2*)c.d_xf;
    uint8_t* d_ridx = (uint8_t*)(d_txy + n);
    SearchRound sr;
    sr.mode = round->mode;
    sr.rot_idx = round->rot_idx;
    sr.n_rots = it.n_rots;
    sr.sd = round->sd;
    for (int k = 0; k < 4; k++) sr.box[k] = round->box[k];
    sr.mean[0] = round->mean[0], sr.mean[1] = round->mean[1];
    sr.seed = round->seed, sr.counter = round->counter;
    cudaStream_t st = e->compute;
    JG_CUDA(e, cudaMemsetAsync(e->d_search, 0xFF, sizeof(unsigned long long), st));
    JG_CUDA(e, cudaMemsetAsync(e->d_search + 1, 0, 3 * sizeof(unsigned long long), st));
    JG_CUDA(e, launch_search_sample(sr, n, d_ridx, d_txy, e->n_sms, st));
    int launches = 3;
    PieceCursor cur;
    Piece p;
    if (!next_piece(e, 1, items.data(), offs, JG_MAX_LAUNCH_QUERIES, cur, p) || p.m != n) return fail(e, JG_ERR_CAPACITY, "search round: batch does not fit one launch");
    JG_CUDA(e, launch_piece(e, plan, p, XfSrc{nullptr, d_ridx, d_txy}, d_filt, d_kinds, JG_MODE_DEFAULT, c.d_out, c.sc, st, &launches));
    JG_CUDA(e, launch_search_reduce(n, c.d_out, d_ridx, d_txy, it.rot_ext, it.n_rots, e->d_search, e->n_sms, st));
    k_search_fetch<<<1, 1, 0, st>>>(e->d_search, d_ridx, d_txy);
    JG_CUDA(e, cudaMemcpyAsync(e->h_search, e->d_search, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    JG_CUDA(e, cudaStreamSynchronize(st));
    e->stat_launches += launches;
    e->stat_queries += (uint64_t)n;
    e->stat_d2h += 32;
    e->search_last_n = n;
    best->n_free = (int64_t)e->h_search[1];
    best->index = -1;
    best->loss = 0.0f, best->tx = 0.0f, best->ty = 0.0f, best->rot_idx = -1;
    if (e->h_search[0] != ~0ull) {
        uint32_t kb = (uint32_t)(e->h_search[0] >> 32);
        kb = (kb & 0x80000000u) ? (kb & 0x7FFFFFFFu) : ~kb;  // inverse of order_key
        std::memcpy(&best->loss, &kb, 4);
        best->index = (int64_t)(uint32_t)e->h_search[0];
        const uint32_t bx = (uint32_t)e->h_search[2], by = (uint32_t)(e->h_search[2] >> 32);
        std::memcpy(&best->tx, &bx, 4);
        std::memcpy(&best->ty, &by, 4);
        best->rot_idx = (int)e->h_search[3];
    }
    return JG_OK;
}

int jg_search_samples(JgEngine* e, int64_t n, uint8_t* rot_idx, float* txy, uint8_t* verdicts) {
    if (!e || n <= 0 || n != e->search_last_n) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    ChunkCtx& c = e->main_ctx;
    const float2* d_txy = (const float2*)c.d_xf;
    const uint8_t* d_ridx = (const uint8_t*)(d_txy + n);
    JG_CUDA(e, cudaStreamSynchronize(e->compute));
    if (txy) JG_CUDA(e, cudaMemcpy(txy, d_txy, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost));
    if (rot_idx) JG_CUDA(e, cudaMemcpy(rot_idx, d_ridx, (size_t)n, cudaMemcpyDeviceToHost));
    if (verdicts) JG_CUDA(e, cudaMemcpy(verdicts, c.d_out, (size_t)n, cudaMemcpyDeviceToHost));
    return JG_OK;
}

int jg_engine_set_batching(JgEngine* e, int64_t chunk_queries, int n_streams) {
    if (!e || chunk_queries < 0 || chunk_queries > JG_MAX_LAUNCH_QUERIES || n_streams < 1 || n_streams > 16) return JG_ERR_INVALID;
    int rc = jg_sync(e);
    if (rc != JG_OK) return rc;
    e->chunk_queries = chunk_queries > 0 ? chunk_queries : (int64_t)1 << 18;
    e->auto_schedule = chunk_queries == 0;  // 0: the per-batch schedule of detect_multi_host (the default)
    e->bufs[0].marks_valid = e->bufs[1].marks_valid = false;  // the set of query streams changes
    while ((int)e->chunks.size() > n_streams) {
        cudaStream_t st = e->chunks.back().st;
        free_ctx(e->chunks.back());
        cudaStreamDestroy(st);
        e->chunks.pop_back();
    }
    while ((int)e->chunks.size() < n_streams) {
        ChunkCtx c;
        JG_CUDA(e, cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        e->chunks.push_back(c);
    }
    for (ChunkCtx& c : e->chunks) {  // buffers are re-sized lazily
        cudaStream_t st = c.st;
        free_ctx(c);
        c.st = st;
    }
    return JG_OK;
}

// Completion point of everything enqueued on this engine so far (the *_async entry points return after enqueueing;
// jg_sync waits for all of it, a fence only for what preceded it). Fences live in a ring of 32.
int jg_fence_record(JgEngine* e, uint64_t* fence) {
    if (!e || !fence) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    if (e->fences.empty()) e->fences.resize(32);
    JgEngine::Fence& f = e->fences[e->next_fence % e->fences.size()];
    std::vector<cudaStream_t> qs{e->compute, e->st_copy};
    for (ChunkCtx& c : e->chunks) qs.push_back(c.st);
    while (f.ev.size() < qs.size()) {
        cudaEvent_t ev;
        JG_CUDA(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        f.ev.push_back(ev);
    }
    for (size_t i = 0; i < qs.size(); i++) JG_CUDA(e, cudaEventRecord(f.ev[i], qs[i]));
    for (size_t i = qs.size(); i < f.ev.size(); i++) JG_CUDA(e, cudaEventRecord(f.ev[i], e->compute));
    f.id = e->next_fence++;
    *fence = f.id;
    return JG_OK;
}
int jg_fence_wait(JgEngine* e, uint64_t fence) {
    if (!e || fence == 0 || fence >= e->next_fence) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JgEngine::Fence& f = e->fences[fence % e->fences.size()];
    if (f.id != fence) return jg_sync(e);  // recycled: everything that old is covered by a full sync
    for (cudaEvent_t ev : f.ev) JG_CUDA(e, cudaEventSynchronize(ev));
    return JG_OK;
}

int jg_engine_set_split_min(JgEngine* e, int64_t min_queries) {
    if (!e || min_queries < 0) return JG_ERR_INVALID;
    e->split_min = min_queries;
    e->plan.split_min = min_queries;
    return JG_OK;
}

int jg_set_dependent_launch(int on) {
    jg::set_dependent_launch(on);
    return JG_OK;
}
int jg_dependent_launch(void) { return jg::dependent_launch(); }

int jg_engine_set_groups(JgEngine* e, int detect_group, int collect_group) {
    auto ok = [](int g) { return g == 0 || g == 4 || g == 8 || g == 16 || g == 32; };
    if (!e || !ok(detect_group) || !ok(collect_group)) return JG_ERR_INVALID;
    e->detect_group = detect_group;
    e->collect_group = collect_group;
    return JG_OK;
}

int jg_engine_set_detect_path(JgEngine* e, int path) {
    if (!e || (path != JG_PATH_AUTO && path != JG_PATH_WALK && path != JG_PATH_FUSED)) return JG_ERR_INVALID;
    int rcs = jg_sync(e);
    if (rcs != JG_OK) return rcs;
    e->force_walk = path == JG_PATH_WALK;
    e->use_fused = path == JG_PATH_FUSED;
    if (e->has_layout && e->d_kinds) e->plan.use_grid = (e->grid_ok_cur && !e->force_walk) ? 1 : 0;
    return JG_OK;
}
int jg_engine_detect_path(const JgEngine* e) { return e && e->plan.use_grid ? (e->use_fused ? 2 : 1) : 0; }

int jg_host_alloc(size_t bytes, void** out) {
    if (!out) return JG_ERR_INVALID;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? JG_OK : JG_ERR_CUDA;
}
int jg_host_alloc_wc(size_t bytes, void** out) {
    if (!out) return JG_ERR_INVALID;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocWriteCombined) == cudaSuccess ? JG_OK : JG_ERR_CUDA;
}
int jg_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? JG_OK : JG_ERR_CUDA; }
int jg_device_alloc(JgEngine* e, size_t bytes, void** out) {
    if (!e || !out) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JG_CUDA(e, cudaMalloc(out, bytes ? bytes : 1));
    return JG_OK;
}
int jg_device_free(JgEngine* e, void* p) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JG_CUDA(e, cudaFree(p));
    return JG_OK;
}
int jg_memcpy_h2d(JgEngine* e, void* d_dst, const void* src, size_t bytes) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JG_CUDA(e, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, e->compute));
    JG_CUDA(e, cudaStreamSynchronize(e->compute));
    return JG_OK;
}
int jg_memcpy_d2h(JgEngine* e, void* dst, const void* d_src, size_t bytes) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    JG_CUDA(e, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, e->compute));
    JG_CUDA(e, cudaStreamSynchronize(e->compute));
    return JG_OK;
}

void jg_compose_batch(const float* rot_tx_ty, int64_t n, float* out4, int n_threads) {
    auto work = [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            const float r = rot_tx_ty[3 * i];
            out4[4 * i + 0] = cosf(r);
            out4[4 * i + 1] = sinf(r);
            out4[4 * i + 2] = rot_tx_ty[3 * i + 1];
            out4[4 * i + 3] = rot_tx_ty[3 * i + 2];
        }
    };
    if (n_threads <= 1 || n < 4096) {
        work(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int k = 0; k < n_threads; k++) th.emplace_back(work, n * k / n_threads, n * (k + 1) / n_threads);
    for (auto& t : th) t.join();
}

int jg_fp32_nofma_peak(JgEngine* e, double* gflops, float* ms_out) {
    if (!e || !gflops) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    const int blocks = e->n_sms * 8, iters = 1 << 15;
    float* d = nullptr;
    JG_CUDA(e, cudaMalloc(&d, (size_t)blocks * 256 * sizeof(float)));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        JG_CUDA(e, cudaEventRecord(e->ev0, e->compute));
        JG_CUDA(e, launch_fp32_nofma(d, blocks, iters, e->compute));
        JG_CUDA(e, cudaEventRecord(e->ev1, e->compute));
        JG_CUDA(e, cudaEventSynchronize(e->ev1));
        float ms = 0;
        JG_CUDA(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    e->timing_valid = false;
    e->stat_launches += 5;
    *gflops = (double)blocks * 256.0 * (double)iters * 8.0 / ((double)best * 1e-3) / 1e9;
    if (ms_out) *ms_out = best;
    return JG_OK;
}

int jg_flush_l2(JgEngine* e) {
    if (!e) return JG_ERR_INVALID;
    JG_CUDA(e, cudaSetDevice(e->device));
    if (!e->d_flush) {
        e->flush_words = (size_t)64 << 20;  // 256 MiB > 126 MB L2
        JG_CUDA(e, cudaMalloc(&e->d_flush, e->flush_words * 4));
    }
    JG_CUDA(e, launch_fill(e->d_flush, e->flush_words, (uint32_t)e->stat_launches, e->n_sms, e->compute));
    e->stat_launches++;
    return JG_OK;
}

// ---- host mirror of CDEngine's update path ------------------------------------------------------------
struct JgCdeBox {
    HostCDE cde;
    std::vector<ImageRange> ranges;  // of the last jg_cde_image_sync
    JgCdeBox(Rect r, JgConfig c) : cde(r, c) {}
};

int jg_cde_create(const float bbox4[4], const JgConfig* cfg, JgCde** out) {
    if (!bbox4 || !cfg || !out) return JG_ERR_INVALID;
    if (cfg->quadtree_depth < 0 || cfg->quadtree_depth > JG_MAX_DEPTH) return JG_ERR_INVALID;
    if (!(bbox4[0] < bbox4[2] && bbox4[1] < bbox4[3])) return JG_ERR_INVALID;
    *out = (JgCde*)new JgCdeBox(Rect{bbox4[0], bbox4[1], bbox4[2], bbox4[3]}, *cfg);
    return JG_OK;
}
void jg_cde_destroy(JgCde* c) { delete (JgCdeBox*)c; }
const char* jg_cde_last_error(const JgCde* c) { return c ? ((const JgCdeBox*)c)->cde.last_error.c_str() : "null cde"; }

int jg_cde_register_hazard(JgCde* c, int kind, int64_t id, int64_t quality, uint64_t pk, const float* verts_xy,
                           int n_verts, const float* poi_xy, int dynamic, uint64_t* hkey) {
    if (!c || !verts_xy || n_verts < 3 || !poi_xy || kind < 0 || kind > 3) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    HostHazard hz;
    hz.kind = kind;
    hz.id = id;
    hz.quality = quality;
    hz.pk = pk;
    hz.dynamic = dynamic != 0;
    hz.xy.assign(verts_xy, verts_xy + 2 * (size_t)n_verts);
    hz.poi_x = poi_xy[0];
    hz.poi_y = poi_xy[1];
    cde.last_error.clear();
    const uint64_t k = cde.register_hazard(std::move(hz));
    if (hkey) *hkey = k;
    return cde.last_error.empty() ? JG_OK : JG_ERR_REFERENCE_PANIC;
}
int jg_cde_place_item(JgCde* c, int64_t item_id, uint64_t pk, const float* item_verts_xy, int n_verts,
                      const float* item_poi_xy, const float transform4[4], uint64_t* hkey) {
    if (!c || !item_verts_xy || n_verts < 3 || !item_poi_xy || !transform4) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    const Xf xf{transform4[0], transform4[1], transform4[2], transform4[3]};
    HostHazard hz;
    hz.kind = JG_HAZ_PLACED_ITEM;
    hz.id = item_id;
    hz.quality = 0;
    hz.pk = pk;
    hz.dynamic = true;
    hz.xy.resize(2 * (size_t)n_verts);
    for (int i = 0; i < n_verts; i++)  // simple_polygon.rs:226-253 transform_from: vertex by vertex
        xf_point(xf, item_verts_xy[2 * i], item_verts_xy[2 * i + 1], hz.xy[2 * i], hz.xy[2 * i + 1]);
    xf_point(xf, item_poi_xy[0], item_poi_xy[1], hz.poi_x, hz.poi_y);  // circle.rs:79-85: centre moves, radius stays
    cde.last_error.clear();
    const uint64_t k = cde.register_hazard(std::move(hz));
    if (hkey) *hkey = k;
    return cde.last_error.empty() ? JG_OK : JG_ERR_REFERENCE_PANIC;
}
int jg_cde_remove_item(JgCde* c, uint64_t pk) {
    if (!c) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    const uint64_t k = cde.hkey_from_pk(pk);
    if (k == 0) return JG_ERR_INVALID;
    return cde.deregister_hazard(k) ? JG_OK : JG_ERR_INVALID;
}
int jg_cde_deregister_hazard(JgCde* c, uint64_t hkey) {
    if (!c) return JG_ERR_INVALID;
    return ((JgCdeBox*)c)->cde.deregister_hazard(hkey) ? JG_OK : JG_ERR_INVALID;
}
int jg_cde_deregister_hazard_by_entity(JgCde* c, int kind, int64_t id, int64_t quality, uint64_t pk) {
    if (!c) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    const uint64_t k = cde.hkey_from_entity(kind, id, quality, pk);
    if (k == 0) {
        cde.last_error = "Cannot deregister hazard that is not registered";  // cd_engine.rs:78
        return JG_ERR_INVALID;
    }
    return cde.deregister_hazard(k) ? JG_OK : JG_ERR_INVALID;
}
int jg_cde_swap_container(JgCde* c, const JgCde* base) {
    if (!c || !base || c == base) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    cde.swap_container(((const JgCdeBox*)base)->cde);
    return cde.last_error.empty() ? JG_OK : JG_ERR_REFERENCE_PANIC;
}
int jg_cde_clone(const JgCde* c, JgCde** out) {
    if (!c || !out) return JG_ERR_INVALID;
    JgCdeBox* b = new JgCdeBox(*(const JgCdeBox*)c);
    b->cde.invalidate_image();
    *out = (JgCde*)b;
    return JG_OK;
}
uint64_t jg_cde_epoch(const JgCde* c) { return c ? ((const JgCdeBox*)c)->cde.epoch : 0; }
int jg_cde_hkey_from_pk(const JgCde* c, uint64_t pk, uint64_t* hkey) {
    if (!c || !hkey) return JG_ERR_INVALID;
    *hkey = ((const JgCdeBox*)c)->cde.hkey_from_pk(pk);
    return JG_OK;
}
int jg_cde_save(JgCde* c, int* snapshot_id) {
    if (!c || !snapshot_id) return JG_ERR_INVALID;
    *snapshot_id = ((JgCdeBox*)c)->cde.save();
    return JG_OK;
}
int jg_cde_restore(JgCde* c, int snapshot_id) {
    if (!c) return JG_ERR_INVALID;
    return ((JgCdeBox*)c)->cde.restore(snapshot_id) ? JG_OK : JG_ERR_INVALID;
}
int jg_cde_slot_capacity(const JgCde* c) { return c ? ((const JgCdeBox*)c)->cde.slot_capacity() : 0; }
int jg_cde_n_hazards(const JgCde* c) { return c ? ((const JgCdeBox*)c)->cde.n_hazards() : 0; }
uint64_t jg_cde_hkey_exterior(const JgCde* c) { return c ? ((const JgCdeBox*)c)->cde.hkey_exterior() : 0; }
int jg_cde_filter_qzones(const JgCde* c, int required_quality, uint32_t* bits, int words) {
    if (!c || !bits || words <= 0) return JG_ERR_INVALID;
    ((const JgCdeBox*)c)->cde.filter_qzones(required_quality, bits, words);
    return JG_OK;
}
int64_t jg_cde_flatten(JgCde* c, void* buf, int64_t cap) {
    if (!c) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    if (cde.hkey_exterior() == 0) {
        cde.last_error = "No exterior hazard registered in the CDE";  // cd_engine.rs:48
        return JG_ERR_INVALID;
    }
    std::vector<uint8_t> blob = cde.flatten();
    if (buf && cap >= (int64_t)blob.size()) std::memcpy(buf, blob.data(), blob.size());
    return (int64_t)blob.size();
}
int jg_cde_image_sync(JgCde* c, const void** image, size_t* bytes, uint64_t* lineage, uint64_t* seq, const uint32_t** ranges_off_len, int* n_ranges) {
    if (!c || !image || !bytes || !lineage || !seq || !ranges_off_len || !n_ranges) return JG_ERR_INVALID;
    JgCdeBox* box = (JgCdeBox*)c;
    if (box->cde.hkey_exterior() == 0) {
        box->cde.last_error = "No exterior hazard registered in the CDE";  // cd_engine.rs:48
        return JG_ERR_INVALID;
    }
    const std::vector<uint8_t>& blob = box->cde.image_sync(box->ranges, *lineage, *seq);
    *image = blob.data();
    *bytes = blob.size();
    *ranges_off_len = box->ranges.empty() ? nullptr : (const uint32_t*)box->ranges.data();
    *n_ranges = (int)box->ranges.size();
    return JG_OK;
}
int jg_cde_sync_to_engine(JgCde* c, JgEngine* e, uint64_t* epoch) {
    if (!c || !e) return JG_ERR_INVALID;
    HostCDE& cde = ((JgCdeBox*)c)->cde;
    if (cde.hkey_exterior() == 0) return fail(e, JG_ERR_INVALID, "No exterior hazard registered in the CDE");
    static const bool incremental = [] { const char* v = getenv("JG_INCREMENTAL"); return !v || atoi(v) != 0; }();
    int rc;
    if (incremental) {  // only what register / deregister touched since the last sync travels
        const void* image;
        size_t bytes;
        uint64_t lineage, seq;
        const uint32_t* ranges;
        int n_ranges;
        rc = jg_cde_image_sync(c, &image, &bytes, &lineage, &seq, &ranges, &n_ranges);
        if (rc != JG_OK) return fail(e, rc, cde.last_error.c_str());
        rc = jg_layout_patch(e, image, bytes, lineage, seq, ranges, n_ranges, cde.epoch);
    } else {
        std::vector<uint8_t> blob = cde.flatten();
        rc = jg_layout_upload(e, blob.data(), blob.size(), cde.epoch);
    }
    if (rc == JG_OK && epoch) *epoch = cde.epoch;
    return rc;
}
#include "engine_comm.inl"

int64_t jg_cde_tree_dump(const JgCde* c, uint32_t* out, int64_t cap) {
    if (!c) return JG_ERR_INVALID;
    std::vector<uint32_t> v = ((const JgCdeBox*)c)->cde.tree_dump();
    if (out)
        for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
}

}  // extern "C"
