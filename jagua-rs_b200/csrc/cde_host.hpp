// Host-side mirror of CDEngine's update path: hazard SlotMap, incremental quadtree maintenance and
// flattening into the SoA blob the kernels read. Index-based arena (no pointer tree): nodes live in
// a deque, the four children of a node are one contiguous block, freed blocks are recycled.
//
// Mirrors (paths under jagua-rs/src/collision_detection):
//   cd_engine.rs:34-149        CDEngine::{new, register_hazard, deregister_hazard_by_key, save, restore}
//   quadtree/qt_node.rs:38-94  QTNode::{register_hazard, deregister_hazard}
//   quadtree/qt_hazard.rs:36-167 QTHazard::{from_root, constrict}
//   quadtree/qt_partial_hazard.rs:15-60 QTHazPartial::{from_entire_shape, from_parent}
//   quadtree/qt_hazard_vec.rs:24-55 QTHazardVec::{add, remove}
#pragma once
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#include "../../include/jagua_cuda.h"
#include "layout_soa.h"
#include "predicates.cuh"

namespace jg {

struct HostHazard {  // hazards/hazard.rs:16-33 (entity + shape + dynamic)
    int kind = JG_HAZ_EXTERIOR;
    int64_t id = 0, quality = 0;
    uint64_t pk = 0;
    bool dynamic = false;
    std::vector<float> xy;  // shape vertices, CCW
    Rect bbox{0, 0, 0, 0};
    float poi_x = 0, poi_y = 0;
    int n_verts() const { return (int)(xy.size() / 2); }
    int scope() const { return kind == JG_HAZ_EXTERIOR ? 0 : 1; }  // hazard.rs:55-62
    Seg edge(int i) const {                                        // simple_polygon.rs:86-93
        int j = (i == n_verts() - 1) ? 0 : i + 1;
        return Seg{xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]};
    }
    bool same_entity(const HostHazard& o) const {
        return kind == o.kind && id == o.id && quality == o.quality && pk == o.pk;
    }
};

enum : uint8_t { PRES_NONE = 0, PRES_PARTIAL = 1, PRES_ENTIRE = 2 };

struct QEntry {  // QTHazard (+ QTHazPartial when presence == PRES_PARTIAL)
    uint32_t slot = 0, version = 0;
    uint8_t presence = PRES_NONE;
    Rect ff_bbox{0, 0, 0, 0};
    std::vector<Seg> edges;
};

struct QNode {
    uint8_t level = 0;
    Rect bbox{0, 0, 0, 0};
    int child_base = -1;  // arena index of child 0; children are child_base .. child_base+3
    int parent = -1;      // arena index of the parent (-1: root)
    bool dead = false;    // lives in a freed child block
    std::vector<QEntry> entries;  // Entire entries first, then Partial
    uint32_t n_active_edges = 0;
};

// A byte range of the device image that changed (HostCDE::image_sync).
struct ImageRange {
    uint32_t off, len;
};

// The host twin of k_grid_build's validity verdict (csrc/detect_grid.cuh) for a NoFilter summary, evaluated on
// a flattened blob: lets a layout upload choose the kernel path without reading anything back from the device.
bool blob_grid_ok(const uint8_t* blob);
bool grid_axes_ok(const float root_bbox[4], int D, float xs[33], float ys[33]);
bool grid_node_ok(const int32_t* node_meta, const float* node_bbox, uint32_t n, uint32_t n_nodes, int D, const float* xs, const float* ys);

class HostCDE {
public:
    HostCDE(Rect bbox, JgConfig cfg);

    uint64_t register_hazard(HostHazard&& hz);
    bool deregister_hazard(uint64_t hkey);
    uint64_t hkey_from_entity(int kind, int64_t id, int64_t quality, uint64_t pk) const;
    // Layout::swap_container (entities/layout.rs:40-50): become a clone of `base` with this CDE's dynamic hazards re-registered
    void swap_container(const HostCDE& base);
    uint64_t hkey_from_pk(uint64_t pk) const;
    int save();
    bool restore(int snapshot_id);

    int slot_capacity() const { return (int)slots_.size(); }
    int n_hazards() const { return n_live_; }
    uint64_t hkey_exterior() const;
    void filter_qzones(int required_quality, uint32_t* bits, int words) const;

    std::vector<uint8_t> flatten() const;
    std::vector<uint32_t> tree_dump() const;

    // ---- incremental device image (BASELINE north_star item 1: "flattened into SoA arrays in HBM and
    // updated incrementally on register/deregister") ----
    // A POSITION-STABLE form of the blob of layout_soa.h: node records at their arena index (freed child blocks stay as dead
    // records, parent = -2), the entries / edges of a node in blocks of their own inside section-sized heaps, hazards by
    // slot, hazard vertices in per-slot blocks. register_hazard / deregister_hazard (and everything built on them: restore,
    // place_item, remove_item) record which nodes and slots they touched; image_sync re-serialises only those and reports the
    // byte ranges that changed since the previous call. `lineage` changes whenever the image had to be rebuilt from scratch
    // (first call, a section outgrew its capacity, fragmentation): the consumer must then take the whole blob.
    const std::vector<uint8_t>& image_sync(std::vector<ImageRange>& ranges, uint64_t& lineage, uint64_t& seq);
    void invalidate_image() { img_.valid = false; }  // (a copy of a CDE must not continue the original's image lineage)

    std::string last_error;
    uint64_t epoch = 0;

private:
    struct Slot {
        uint32_t version = 0;  // odd = occupied (slotmap 1.0)
        uint32_t next_free = UINT32_MAX;
        bool live = false;
        HostHazard hz;
    };
    static uint64_t make_key(uint32_t idx, uint32_t version) { return ((uint64_t)version << 32) | idx; }

    int alloc_children(int parent, const Rect& bbox, uint8_t child_level);
    void touch_node(int node);
    void touch_slot(uint32_t slot);
    void free_children(int child_base);
    void node_register(int node, QEntry&& e);
    void node_deregister(int node, uint32_t slot, uint32_t version);
    void constrict_into_children(int node, const QEntry& e);
    void constrict(const QEntry& e, const Rect quads[4], QEntry out[4]) const;
    static void entry_insert(QNode& n, QEntry&& e);
    void dump_node(int node, std::vector<uint32_t>& out) const;

    JgConfig cfg_;
    std::deque<QNode> nodes_;       // nodes_[0] is the root
    std::vector<int> free_blocks_;  // recycled child blocks
    std::vector<Slot> slots_;
    uint32_t free_head_ = UINT32_MAX;
    int n_live_ = 0;
    std::vector<std::vector<HostHazard>> snapshots_;

    // image state (cde_image.cpp)
    struct Block {
        uint32_t off = 0, cap = 0;  // in elements of the section
    };
    struct Heap {  // element heap of one section: bump pointer + free lists by exact block size
        uint32_t top = 0, cap = 0, live = 0;
        std::vector<std::vector<uint32_t>> free_by_size;
        bool alloc(uint32_t n, Block& b);
        void release(Block& b);
    };
    struct Image {
        std::vector<uint8_t> blob;
        JgLayoutHeader hdr{};
        uint32_t cap_nodes = 0, cap_slots = 0;
        Heap entries, edges, verts;
        std::vector<Block> node_entries, node_edges, slot_verts;
        std::vector<ImageRange> dirty;
        uint64_t lineage = 0, seq = 0;
        bool valid = false;
        // the grid premise, kept up to date node by node (the twin of k_grid_build's verdict, see blob_grid_ok)
        bool axes_ok = false;
        float xs[33], ys[33];
        std::vector<uint8_t> node_bad;
        int n_bad = 0;
    } img_;
    std::vector<int> dirty_nodes_;
    std::vector<uint8_t> node_dirty_;
    std::vector<uint32_t> dirty_slots_;
    static uint64_t next_lineage_;
    void image_rebuild();
    bool image_write_node(int node);
    void image_check_node(int node);
    bool image_write_slot(uint32_t slot);
    void image_mark(uint32_t off, uint32_t len);
    void image_write_header();
};

}  // namespace jg
