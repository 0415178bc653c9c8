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
    std::vector<QEntry> entries;  // Entire entries first, then Partial
    uint32_t n_active_edges = 0;
};

// The host twin of k_grid_build's validity verdict (csrc/detect_grid.cuh) for a NoFilter summary, evaluated on
// a flattened blob: lets a layout upload choose the kernel path without reading anything back from the device.
bool blob_grid_ok(const uint8_t* blob);

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

    int alloc_children(const Rect& bbox, uint8_t child_level);
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
};

}  // namespace jg
