// Layout broadcast over NCCL (BASELINE north_star: "NCCL over NVLink is used only to broadcast layout updates"). Included by
// engine.cu inside its extern "C" block (it needs JgEngine's internals). NCCL is resolved at run time with dlopen: a process
// that already holds a libnccl.so.2 (torch's bundled one) shares it, others load the system library; an engine that never
// calls jg_engine_comm_init needs no NCCL at all.
//
// Protocol of one update (every rank of the communicator calls jg_layout_broadcast, in the same order):
//   root      its own layout: the ordinary stream-ordered jg_layout_upload. For the others: blob -> pinned staging ->
//             d_bcast (H2D) -> ncclBroadcast of the FIXED capacity of the exchange buffer, all on st_bc. No host wait.
//   others    ncclBroadcast into d_bcast -> k_bcast_check (parses the header ON THE DEVICE and runs the structural checks
//             of k_blob_validate) -> 132-byte read-back {header, violations} -> ONE host wait on st_bc -> the stream-ordered
//             upload from device memory (D2D on st_up, summaries), which marks d_bcast as consumed with an event.
// The count is fixed (the capacity agreed at jg_engine_comm_init) because ncclBroadcast needs the same count on every rank
// and the receivers learn the blob's size only from its header: one collective, one host wait per update.
#include <dlfcn.h>

namespace {
typedef struct { char internal[128]; } JgNcclId;
typedef int (*fn_nccl_get_unique_id)(JgNcclId*);
typedef int (*fn_nccl_comm_init_rank)(void**, int, JgNcclId, int);
typedef int (*fn_nccl_broadcast)(const void*, void*, size_t, int /*ncclDataType_t*/, int, void*, cudaStream_t);
typedef int (*fn_nccl_comm_destroy)(void*);
typedef const char* (*fn_nccl_get_error_string)(int);
struct NcclApi {
    void* lib = nullptr;
    fn_nccl_get_unique_id get_unique_id = nullptr;
    fn_nccl_comm_init_rank comm_init_rank = nullptr;
    fn_nccl_broadcast broadcast = nullptr;
    fn_nccl_comm_destroy comm_destroy = nullptr;
    fn_nccl_get_error_string error_string = nullptr;
    std::string why;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            a.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return a;
        }
        a.get_unique_id = (fn_nccl_get_unique_id)dlsym(a.lib, "ncclGetUniqueId");
        a.comm_init_rank = (fn_nccl_comm_init_rank)dlsym(a.lib, "ncclCommInitRank");
        a.broadcast = (fn_nccl_broadcast)dlsym(a.lib, "ncclBroadcast");
        a.comm_destroy = (fn_nccl_comm_destroy)dlsym(a.lib, "ncclCommDestroy");
        a.error_string = (fn_nccl_get_error_string)dlsym(a.lib, "ncclGetErrorString");
        if (!a.get_unique_id || !a.comm_init_rank || !a.broadcast || !a.comm_destroy) {
            a.why = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclBroadcast / ncclCommDestroy";
            a.lib = nullptr;
        }
        return a;
    }();
    return api;
}
int nccl_fail(JgEngine* e, const char* what, int rc) {
    NcclApi& a = nccl_api();
    e->err = std::string(what) + ": " + (a.error_string ? a.error_string(rc) : "NCCL error") + " (" + std::to_string(rc) + ")";
    return JG_ERR_CUDA;
}
}  // namespace

// header parsed on the device + the checks of k_blob_validate; out[0] = violations (header or structure)
__global__ void k_bcast_check(const unsigned char* __restrict__ blob, size_t cap, JgLayoutHeader* __restrict__ hdr_out, int* __restrict__ bad) {
    __shared__ JgLayoutHeader h;
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        h = *(const JgLayoutHeader*)blob;
        int b = 0;
        if (h.magic != JG_LAYOUT_MAGIC || h.version != JG_LAYOUT_VERSION || h.total_bytes > cap || h.n_nodes == 0 || h.n_nodes > JG_MAX_NODES ||
            h.root_level > JG_MAX_DEPTH)
            b = 1;
        const uint32_t off[11] = {h.off_node_meta, h.off_node_bbox, h.off_node_nactive, h.off_entry_meta, h.off_entry_ffbbox, h.off_edges,
                                  h.off_haz_meta, h.off_haz_bbox, h.off_haz_poi, h.off_haz_verts, h.off_haz_version};
        const uint32_t cnt[11] = {h.n_nodes, h.n_nodes, h.n_nodes, h.n_entries, h.n_entries, h.n_edges, h.n_slots, h.n_slots, h.n_slots, h.n_haz_verts, h.n_slots};
        const uint32_t sz[11] = {16, 16, 4, 16, 16, 16, 16, 16, 16, 8, 4};
        for (int k = 0; k < 11 && !b; k++)
            if ((off[k] & 15u) || (unsigned long long)off[k] + (unsigned long long)cnt[k] * sz[k] > h.total_bytes) b = 1;
        s_bad = b;
        if (blockIdx.x == 0) *hdr_out = h;
    }
    __syncthreads();
    if (s_bad) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(bad, 1);
        return;
    }
    const int4* node_meta = (const int4*)(blob + h.off_node_meta);
    const float4* node_bbox = (const float4*)(blob + h.off_node_bbox);
    const int4* entry_meta = (const int4*)(blob + h.off_entry_meta);
    const int4* haz_meta = (const int4*)(blob + h.off_haz_meta);
    const int n_nodes = (int)h.n_nodes, n_entries = (int)h.n_entries, n_edges = (int)h.n_edges, n_slots = (int)h.n_slots, n_hv = (int)h.n_haz_verts;
    uint32_t n = h.n_nodes > h.n_entries ? h.n_nodes : h.n_entries;
    if (h.n_slots > n) n = h.n_slots;
    bool b = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)n; i += gridDim.x * blockDim.x) {
        if (i < n_nodes) {
            const int4 m = node_meta[i];
            const int n_ent = (m.z & 0xFFFF) + ((m.z >> 16) & 0xFFFF);
            if (m.x >= 0 && ((m.x & 3) != 1 || m.x + 4 > n_nodes)) b = true;
            if (m.y < 0 || m.y + n_ent > n_entries) b = true;
            if (m.w >= n_nodes || (i > 0 && m.w < 0 && m.w != -2)) b = true;
            const float4 r = node_bbox[i];
            if (!(fabsf(r.x) < 1e30f && fabsf(r.y) < 1e30f && fabsf(r.z) < 1e30f && fabsf(r.w) < 1e30f && r.x <= r.z && r.y <= r.w)) b = true;
        }
        if (i < n_entries) {
            const int4 em = entry_meta[i];
            if (em.x < 0 || em.x >= n_slots || em.y < 0 || em.z < 0 || em.y + em.z > n_edges) b = true;
        }
        if (i < n_slots) {
            const int4 hm = haz_meta[i];
            if (hm.y < 0 || hm.z < 0 || hm.y + hm.z > n_hv) b = true;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && h.exterior_slot >= h.n_slots) b = true;
    if (b) atomicAdd(bad, 1);
}

int jg_nccl_unique_id(uint8_t id[128]) {
    if (!id) return JG_ERR_INVALID;
    NcclApi& a = nccl_api();
    if (!a.lib) return JG_ERR_CUDA;
    JgNcclId u;
    if (a.get_unique_id(&u) != 0) return JG_ERR_CUDA;
    std::memcpy(id, u.internal, 128);
    return JG_OK;
}

int jg_engine_comm_destroy(JgEngine* e) {
    if (!e) return JG_ERR_INVALID;
    cudaSetDevice(e->device);
    if (e->st_bc) cudaStreamSynchronize(e->st_bc);
    if (e->nccl_comm && nccl_api().comm_destroy) nccl_api().comm_destroy(e->nccl_comm);
    e->nccl_comm = nullptr;
    e->comm_rank = -1, e->comm_world = 0;
    if (e->d_bcast) cudaFree(e->d_bcast);
    if (e->h_bcast_stage) cudaFreeHost(e->h_bcast_stage);
    if (e->h_bcast_hdr) cudaFreeHost(e->h_bcast_hdr);
    if (e->ev_bc_consumed) cudaEventDestroy(e->ev_bc_consumed);
    if (e->st_bc) cudaStreamDestroy(e->st_bc);
    e->d_bcast = nullptr, e->h_bcast_stage = nullptr, e->h_bcast_hdr = nullptr, e->ev_bc_consumed = nullptr, e->st_bc = nullptr;
    e->bcast_cap = 0;
    return JG_OK;
}

int jg_engine_comm_init(JgEngine* e, const uint8_t id[128], int rank, int world, size_t max_blob_bytes) {
    if (!e || !id || world < 1 || rank < 0 || rank >= world) return JG_ERR_INVALID;
    NcclApi& a = nccl_api();
    if (!a.lib) return fail(e, JG_ERR_CUDA, a.why.c_str());
    JG_CUDA(e, cudaSetDevice(e->device));
    jg_engine_comm_destroy(e);
    if (max_blob_bytes < 4096) max_blob_bytes = (size_t)1 << 20;
    e->bcast_cap = (max_blob_bytes + 255) & ~(size_t)255;
    JG_CUDA(e, cudaStreamCreateWithFlags(&e->st_bc, cudaStreamNonBlocking));
    JG_CUDA(e, cudaMalloc(&e->d_bcast, e->bcast_cap));
    JG_CUDA(e, cudaMemset(e->d_bcast, 0, e->bcast_cap));
    JG_CUDA(e, cudaHostAlloc(&e->h_bcast_stage, e->bcast_cap, cudaHostAllocDefault));
    JG_CUDA(e, cudaHostAlloc((void**)&e->h_bcast_hdr, sizeof(JgLayoutHeader) + 16, cudaHostAllocDefault));
    JG_CUDA(e, cudaEventCreateWithFlags(&e->ev_bc_consumed, cudaEventDisableTiming));
    JG_CUDA(e, cudaEventRecord(e->ev_bc_consumed, e->st_bc));
    JgNcclId u;
    std::memcpy(u.internal, id, 128);
    const int rc = a.comm_init_rank(&e->nccl_comm, world, u, rank);
    if (rc != 0) {
        e->nccl_comm = nullptr;
        return nccl_fail(e, "ncclCommInitRank", rc);
    }
    e->comm_rank = rank, e->comm_world = world;
    return JG_OK;
}

int jg_layout_broadcast(JgEngine* e, int root, const void* blob, size_t bytes, uint64_t epoch) {
    if (!e || !e->nccl_comm) return e ? fail(e, JG_ERR_INVALID, "jg_layout_broadcast: no communicator (jg_engine_comm_init)") : JG_ERR_INVALID;
    if (root < 0 || root >= e->comm_world) return JG_ERR_INVALID;
    NcclApi& a = nccl_api();
    JG_CUDA(e, cudaSetDevice(e->device));
    const bool is_root = e->comm_rank == root;
    int* d_bad = e->d_blob_bad + 1;
    if (is_root) {
        if (!blob || bytes < sizeof(JgLayoutHeader)) return fail(e, JG_ERR_INVALID, "jg_layout_broadcast: the root needs the flattened layout");
        if (bytes > e->bcast_cap) return fail(e, JG_ERR_CAPACITY, "jg_layout_broadcast: the blob exceeds the capacity agreed at jg_engine_comm_init");
        // (the epoch travels in the header's spare word: its low 32 bits)
        int rc = jg_layout_upload(e, blob, bytes, epoch);  // the root's own layout: the ordinary stream-ordered upload
        if (rc != JG_OK) return rc;
        JG_CUDA(e, cudaStreamSynchronize(e->st_bc));  // the previous broadcast has left the staging buffer
        std::memcpy(e->h_bcast_stage, blob, bytes);
        ((JgLayoutHeader*)e->h_bcast_stage)->reserved1 = (uint32_t)epoch;
        JG_CUDA(e, cudaMemcpyAsync(e->d_bcast, e->h_bcast_stage, bytes, cudaMemcpyHostToDevice, e->st_bc));
        const int nrc = a.broadcast(e->d_bcast, e->d_bcast, e->bcast_cap, 0 /*ncclInt8*/, root, e->nccl_comm, e->st_bc);
        if (nrc != 0) return nccl_fail(e, "ncclBroadcast", nrc);
        e->stat_h2d += bytes;
        return JG_OK;
    }
    // receivers: d_bcast is overwritten only after the previous blob has been copied out of it
    JG_CUDA(e, cudaStreamWaitEvent(e->st_bc, e->ev_bc_consumed, 0));
    const int nrc = a.broadcast(e->d_bcast, e->d_bcast, e->bcast_cap, 0 /*ncclInt8*/, root, e->nccl_comm, e->st_bc);
    if (nrc != 0) return nccl_fail(e, "ncclBroadcast", nrc);
    JG_CUDA(e, cudaMemsetAsync(d_bad, 0, sizeof(int), e->st_bc));
    // the device's copy of the parsed header lives in a small buffer of its own (the blob may fill the exchange buffer)
    if (!e->d_hdr_scratch) JG_CUDA(e, cudaMalloc(&e->d_hdr_scratch, sizeof(JgLayoutHeader)));
    k_bcast_check<<<32, 256, 0, e->st_bc>>>((const unsigned char*)e->d_bcast, e->bcast_cap, (JgLayoutHeader*)e->d_hdr_scratch, d_bad);
    JG_CUDA(e, cudaGetLastError());
    JG_CUDA(e, cudaMemcpyAsync(e->h_bcast_hdr, e->d_hdr_scratch, sizeof(JgLayoutHeader), cudaMemcpyDeviceToHost, e->st_bc));
    JG_CUDA(e, cudaMemcpyAsync((char*)e->h_bcast_hdr + sizeof(JgLayoutHeader), d_bad, sizeof(int), cudaMemcpyDeviceToHost, e->st_bc));
    JG_CUDA(e, cudaStreamSynchronize(e->st_bc));  // the one host wait of an update on a receiving rank
    e->stat_launches += 1;
    int n_bad = 0;
    std::memcpy(&n_bad, (char*)e->h_bcast_hdr + sizeof(JgLayoutHeader), sizeof(int));
    if (n_bad) {
        JG_CUDA(e, cudaEventRecord(e->ev_bc_consumed, e->st_bc));
        return fail(e, JG_ERR_INVALID, "jg_layout_broadcast: the received blob is malformed (header or structure)");
    }
    const JgLayoutHeader h = *e->h_bcast_hdr;
    (void)epoch;
    return upload_common(e, e->d_bcast, h.total_bytes, (uint64_t)h.reserved1, true, &h);
}

int jg_cde_broadcast(JgCde* c, JgEngine* e, int root, uint64_t* epoch) {
    if (!e) return JG_ERR_INVALID;
    if (e->comm_rank == root) {
        if (!c) return fail(e, JG_ERR_INVALID, "jg_cde_broadcast: the root needs the host CDE");
        HostCDE& cde = ((JgCdeBox*)c)->cde;
        if (cde.hkey_exterior() == 0) return fail(e, JG_ERR_INVALID, "No exterior hazard registered in the CDE");
        std::vector<uint8_t> blob = cde.flatten();
        const int rc = jg_layout_broadcast(e, root, blob.data(), blob.size(), cde.epoch);
        if (rc == JG_OK && epoch) *epoch = cde.epoch;
        return rc;
    }
    const int rc = jg_layout_broadcast(e, root, nullptr, 0, 0);
    if (rc == JG_OK && epoch) *epoch = e->epoch;
    return rc;
}
