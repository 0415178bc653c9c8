// In-process multi-device pool (include/jagua_cuda.h, jg_pool_*): one JgEngine per CUDA device, the layout / items /
// filters replicated on every device, ONE host batch split into contiguous shards (SURVEY section 8(e):
// [g n/G, (g+1) n/G)), every device copying its shard's results straight into its own slice of the caller's result
// array. No collective on the query path. One persistent host thread per device does that device's enqueueing, so the
// per-launch driver work of the devices overlaps instead of queueing up in the caller's thread.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/jagua_cuda.h"

namespace {

struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv, idle_cv;
    std::deque<std::function<int()>> q;
    bool stop = false;
    std::atomic<int> pending{0};  // tasks queued (what a spinning worker polls)
    int busy = 0;      // tasks queued or running
    int first_rc = 0;  // first failure since the last jg_pool_sync
    void run() {
        for (;;) {
            std::function<int()> fn;
            {
                // a short spin before blocking: batches usually follow each other within microseconds, and waking a blocked
                // thread costs 10-50 us - a third of a small shard's device time
                const auto t0 = std::chrono::steady_clock::now();
                while (pending.load(std::memory_order_acquire) == 0 && std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200))
                    std::this_thread::yield();
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                fn = std::move(q.front());
                q.pop_front();
                pending.fetch_sub(1, std::memory_order_release);
            }
            const int rc = fn();
            {
                std::lock_guard<std::mutex> lk(m);
                if (rc != JG_OK && first_rc == JG_OK) first_rc = rc;
                busy--;
                if (busy == 0) idle_cv.notify_all();
            }
        }
    }
    void push(std::function<int()> fn) {
        {
            std::lock_guard<std::mutex> lk(m);
            q.push_back(std::move(fn));
            busy++;
            pending.fetch_add(1, std::memory_order_release);
        }
        cv.notify_one();
    }
    int drain() {  // waits until every queued task has run; returns and clears the first failure
        std::unique_lock<std::mutex> lk(m);
        idle_cv.wait(lk, [&] { return busy == 0; });
        const int rc = first_rc;
        first_rc = JG_OK;
        return rc;
    }
};

// contiguous shard [lo, hi) of an n-query batch for part k of `parts` (sizes differ by at most one)
inline void shard_bounds(int64_t n, int k, int parts, int64_t& lo, int64_t& hi) {
    const int64_t base = n / parts, rem = n % parts;
    lo = k * base + (k < rem ? k : rem);
    hi = lo + base + (k < rem ? 1 : 0);
}

struct Shard {
    int64_t lo = 0, hi = 0;
    std::vector<int> ids;
    std::vector<int64_t> offs;  // relative to lo
};

// the part of a segmented batch that falls into [lo, hi)
Shard make_shard(int n_seg, const int* item_ids, const int64_t* seg_offsets, int64_t lo, int64_t hi) {
    Shard s;
    s.lo = lo, s.hi = hi;
    s.offs.push_back(0);
    for (int k = 0; k < n_seg; k++) {
        const int64_t a = seg_offsets[k] > lo ? seg_offsets[k] : lo, b = seg_offsets[k + 1] < hi ? seg_offsets[k + 1] : hi;
        if (b > a) {
            s.ids.push_back(item_ids[k]);
            s.offs.push_back(s.offs.back() + (b - a));
        }
    }
    return s;
}

}  // namespace

struct JgPool {
    std::vector<JgEngine*> eng;
    std::vector<int> devices;
    std::vector<Worker*> workers;
    std::string err;
    int64_t min_shard = 1 << 15;  // queries per device below which a batch is spread over fewer devices
};

static int pool_fail(JgPool* p, int code, const std::string& msg) {
    p->err = msg;
    return code;
}

// runs fn(k) on the worker of every device in [0, n_active) and waits for all of them
static int pool_run_all(JgPool* p, int n_active, const std::function<int(int)>& fn) {
    for (int k = 0; k < n_active; k++) p->workers[k]->push([fn, k] { return fn(k); });
    int rc = JG_OK;
    for (int k = 0; k < n_active; k++) {
        const int r = p->workers[k]->drain();
        if (r != JG_OK && rc == JG_OK) {
            rc = r;
            p->err = "device " + std::to_string(p->devices[k]) + ": " + jg_last_error(p->eng[k]);
        }
    }
    return rc;
}

// waits for the queued tasks of the first n_active workers; first failure wins
static int pool_drain(JgPool* p, int n_active) {
    int rc = JG_OK;
    for (int k = 0; k < n_active; k++) {
        const int r = p->workers[k]->drain();
        if (r != JG_OK && rc == JG_OK) {
            rc = r;
            p->err = "device " + std::to_string(p->devices[k]) + ": " + jg_last_error(p->eng[k]);
        }
    }
    return rc;
}
static int active_devices_of(int64_t n, int n_devices, int64_t min_shard) {
    int64_t a = (n + min_shard - 1) / min_shard;
    if (a < 1) a = 1;
    return a < (int64_t)n_devices ? (int)a : n_devices;
}
static int active_devices(const JgPool* p, int64_t n) { return active_devices_of(n, (int)p->eng.size(), p->min_shard); }

extern "C" {

// The pool's sharding rule as a pure function (no device needed): device k of n_devices resolves queries [lo, hi) of an
// n-query batch; devices beyond the active ones (a batch is not cut below min_shard queries per device) get lo == hi.
int jg_pool_shard_bounds(int64_t n, int k, int n_devices, int64_t min_shard, int64_t* lo, int64_t* hi) {
    if (n < 0 || n_devices < 1 || k < 0 || k >= n_devices || min_shard < 1 || !lo || !hi) return JG_ERR_INVALID;
    const int act = active_devices_of(n, n_devices, min_shard);
    if (k >= act) {
        *lo = *hi = n;
        return JG_OK;
    }
    shard_bounds(n, k, act, *lo, *hi);
    return JG_OK;
}

int jg_pool_create(const int* devices, int n_devices, const JgConfig* cfg, JgPool** out) {
    if (!devices || n_devices < 1 || n_devices > 64 || !cfg || !out) return JG_ERR_INVALID;
    *out = nullptr;
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return JG_ERR_INVALID;
    JgPool* p = new JgPool();
    for (int k = 0; k < n_devices; k++) {
        JgEngine* e = nullptr;
        const int rc = jg_engine_create(devices[k], cfg, &e);
        if (rc != JG_OK) {
            for (JgEngine* x : p->eng) jg_engine_destroy(x);
            delete p;
            return rc;
        }
        p->eng.push_back(e);
        p->devices.push_back(devices[k]);
    }
    for (int k = 0; k < n_devices; k++) {
        Worker* w = new Worker();
        w->th = std::thread([w] { w->run(); });
        p->workers.push_back(w);
    }
    *out = p;
    return JG_OK;
}

void jg_pool_destroy(JgPool* p) {
    if (!p) return;
    for (Worker* w : p->workers) {
        w->drain();
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->stop = true;
        }
        w->cv.notify_all();
        w->th.join();
        delete w;
    }
    for (JgEngine* e : p->eng) jg_engine_destroy(e);
    delete p;
}

int jg_pool_size(const JgPool* p) { return p ? (int)p->eng.size() : 0; }
JgEngine* jg_pool_engine(JgPool* p, int k) { return (p && k >= 0 && k < (int)p->eng.size()) ? p->eng[k] : nullptr; }
const char* jg_pool_last_error(const JgPool* p) { return p ? p->err.c_str() : "null pool"; }

int jg_pool_set_min_shard(JgPool* p, int64_t min_queries_per_device) {
    if (!p || min_queries_per_device < 1) return JG_ERR_INVALID;
    p->min_shard = min_queries_per_device;
    return JG_OK;
}

int jg_pool_item_upload(JgPool* p, int item_id, const float* verts_xy, int n_verts, const float* poi_xy, const float* poles_xyr,
                        int n_poles, const float* piers, int n_piers) {
    if (!p) return JG_ERR_INVALID;
    return pool_run_all(p, (int)p->eng.size(), [=](int k) {
        return jg_item_upload(p->eng[k], item_id, verts_xy, n_verts, poi_xy, poles_xyr, n_poles, piers, n_piers);
    });
}

int jg_pool_item_set_rotations(JgPool* p, int item_id, const float* rot_cos_sin, int n_rot) {
    if (!p) return JG_ERR_INVALID;
    return pool_run_all(p, (int)p->eng.size(), [=](int k) { return jg_item_set_rotations(p->eng[k], item_id, rot_cos_sin, n_rot); });
}

int jg_pool_filter_upload(JgPool* p, int filter_id, const uint32_t* bits, int words) {
    if (!p) return JG_ERR_INVALID;
    return pool_run_all(p, (int)p->eng.size(), [=](int k) { return jg_filter_upload(p->eng[k], filter_id, bits, words); });
}

// The layout is small (tens to hundreds of KB), so it is replicated: every device's worker stages the blob in its own
// pinned buffer and copies it on its own upload stream (stream-ordered: queries in flight finish on the old layout).
// Returns once every device has taken its copy of the blob; the caller's buffer is free again.
int jg_pool_layout_upload(JgPool* p, const void* blob, size_t bytes, uint64_t epoch) {
    if (!p || !blob) return JG_ERR_INVALID;
    return pool_run_all(p, (int)p->eng.size(), [=](int k) { return jg_layout_upload(p->eng[k], blob, bytes, epoch); });
}

int jg_pool_cde_sync(JgCde* c, JgPool* p, uint64_t* epoch) {
    if (!c || !p) return JG_ERR_INVALID;
    // the host mirror's incremental image: synced ONCE, the same ranges go to every device
    const void* image;
    size_t bytes;
    uint64_t lineage, seq;
    const uint32_t* ranges;
    int n_ranges;
    int rc = jg_cde_image_sync(c, &image, &bytes, &lineage, &seq, &ranges, &n_ranges);
    if (rc != JG_OK) return pool_fail(p, rc, jg_cde_last_error(c));
    const uint64_t ep = jg_cde_epoch(c);
    rc = pool_run_all(p, (int)p->eng.size(), [=](int k) { return jg_layout_patch(p->eng[k], image, bytes, lineage, seq, ranges, n_ranges, ep); });
    if (rc == JG_OK && epoch) *epoch = ep;
    return rc;
}

// One host batch, sharded. kind 0: composed 16 B records, 1: indexed 9 B form, 2: lattice 5 B form.
static int pool_detect(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                       const uint8_t* rot_idx, const float* txy, const JgLattice* lat, const uint16_t* ij, int filter_id, int mode,
                       uint8_t* out, bool wait) {
    if (!p || n_seg <= 0 || !item_ids || !seg_offsets || seg_offsets[0] != 0) return JG_ERR_INVALID;
    const int64_t n = seg_offsets[n_seg];
    if (n < 0) return JG_ERR_INVALID;
    if (n == 0) return JG_OK;
    const int act = active_devices(p, n);
    const JgLattice lat_copy = lat ? *lat : JgLattice{};
    for (int k = 0; k < act; k++) {
        int64_t lo, hi;
        shard_bounds(n, k, act, lo, hi);
        // the segment table is copied into the task: the caller's arrays need not outlive this call
        Shard s = make_shard(n_seg, item_ids, seg_offsets, lo, hi);
        JgEngine* e = p->eng[k];
        p->workers[k]->push([=] {
            if (s.ids.empty()) return (int)JG_OK;
            const int ns = (int)s.ids.size();
            int rc;
            if (transforms) rc = jg_detect_multi_async(e, ns, s.ids.data(), s.offs.data(), transforms + 4 * s.lo, filter_id, mode, out + s.lo);
            else if (ij) rc = jg_detect_multi_lattice_async(e, ns, s.ids.data(), s.offs.data(), &lat_copy, rot_idx + s.lo, ij + 2 * s.lo, filter_id, mode, out + s.lo);
            else rc = jg_detect_multi_indexed_async(e, ns, s.ids.data(), s.offs.data(), rot_idx + s.lo, txy + 2 * s.lo, filter_id, mode, out + s.lo);
            if (rc == JG_OK && wait) rc = jg_sync(e);  // the blocking form waits on the worker: one hand-over instead of two
            return rc;
        });
    }
    if (!wait) return JG_OK;
    return pool_drain(p, act);
}

int jg_pool_detect_multi(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int filter_id,
                         int mode, uint8_t* out) {
    if (!transforms || !out) return JG_ERR_INVALID;
    return pool_detect(p, n_seg, item_ids, seg_offsets, transforms, nullptr, nullptr, nullptr, nullptr, filter_id, mode, out, true);
}
int jg_pool_detect_multi_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms,
                               int filter_id, int mode, uint8_t* out) {
    if (!transforms || !out) return JG_ERR_INVALID;
    return pool_detect(p, n_seg, item_ids, seg_offsets, transforms, nullptr, nullptr, nullptr, nullptr, filter_id, mode, out, false);
}
int jg_pool_detect_multi_indexed(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                                 const float* txy, int filter_id, int mode, uint8_t* out) {
    if (!rot_idx || !txy || !out) return JG_ERR_INVALID;
    return pool_detect(p, n_seg, item_ids, seg_offsets, nullptr, rot_idx, txy, nullptr, nullptr, filter_id, mode, out, true);
}
int jg_pool_detect_multi_indexed_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const uint8_t* rot_idx,
                                       const float* txy, int filter_id, int mode, uint8_t* out) {
    if (!rot_idx || !txy || !out) return JG_ERR_INVALID;
    return pool_detect(p, n_seg, item_ids, seg_offsets, nullptr, rot_idx, txy, nullptr, nullptr, filter_id, mode, out, false);
}
int jg_pool_detect_multi_lattice(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                                 const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out) {
    if (!lattice || !rot_idx || !ij || !out) return JG_ERR_INVALID;
    return pool_detect(p, n_seg, item_ids, seg_offsets, nullptr, rot_idx, nullptr, lattice, ij, filter_id, mode, out, true);
}
int jg_pool_detect_multi_lattice_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const JgLattice* lattice,
                                       const uint8_t* rot_idx, const uint16_t* ij, int filter_id, int mode, uint8_t* out) {
    if (!lattice || !rot_idx || !ij || !out) return JG_ERR_INVALID;
    return pool_detect(p, n_seg, item_ids, seg_offsets, nullptr, rot_idx, nullptr, lattice, ij, filter_id, mode, out, false);
}

static int pool_collect(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                        const uint32_t* seed_bits, uint32_t* out_bits, int words, bool wait) {
    if (!p || n_seg <= 0 || !item_ids || !seg_offsets || seg_offsets[0] != 0 || !transforms || !out_bits || words <= 0) return JG_ERR_INVALID;
    const int64_t n = seg_offsets[n_seg];
    if (n < 0) return JG_ERR_INVALID;
    if (n == 0) return JG_OK;
    const int act = active_devices(p, n);
    for (int k = 0; k < act; k++) {
        int64_t lo, hi;
        shard_bounds(n, k, act, lo, hi);
        Shard s = make_shard(n_seg, item_ids, seg_offsets, lo, hi);
        JgEngine* e = p->eng[k];
        p->workers[k]->push([=] {
            if (s.ids.empty()) return (int)JG_OK;
            return jg_collect_multi_async(e, (int)s.ids.size(), s.ids.data(), s.offs.data(), transforms + 4 * s.lo, mode,
                                          seed_bits ? seed_bits + (size_t)s.lo * words : nullptr, out_bits + (size_t)s.lo * words, words);
        });
    }
    return wait ? jg_pool_sync(p) : (int)JG_OK;
}
int jg_pool_collect_multi(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                          const uint32_t* seed_bits, uint32_t* out_bits, int words) {
    return pool_collect(p, n_seg, item_ids, seg_offsets, transforms, mode, seed_bits, out_bits, words, true);
}
int jg_pool_collect_multi_async(JgPool* p, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* transforms, int mode,
                                const uint32_t* seed_bits, uint32_t* out_bits, int words) {
    return pool_collect(p, n_seg, item_ids, seg_offsets, transforms, mode, seed_bits, out_bits, words, false);
}

// waits for every device: first for its worker's queued enqueues, then (on the workers, in parallel) for the device
int jg_pool_sync(JgPool* p) {
    if (!p) return JG_ERR_INVALID;
    int rc = JG_OK;
    for (size_t k = 0; k < p->eng.size(); k++) {
        JgEngine* e = p->eng[k];
        p->workers[k]->push([e] { return jg_sync(e); });
    }
    for (size_t k = 0; k < p->eng.size(); k++) {
        const int r = p->workers[k]->drain();
        if (r != JG_OK && rc == JG_OK) {
            rc = r;
            p->err = "device " + std::to_string(p->devices[k]) + ": " + jg_last_error(p->eng[k]);
        }
    }
    return rc;
}

}  // extern "C"
