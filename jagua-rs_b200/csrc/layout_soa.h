// Flattened layout ("blob"): the QTNode quadtree, its QTHazard entries and the hazard shapes as
// structure-of-arrays sections of one contiguous buffer. This is what lives in HBM (and, when it
// fits, is staged into shared memory by the kernels). Every section starts 16-byte aligned so the
// kernels can use 128-bit loads.
//
// Reference structures being flattened (paths under jagua-rs/src/collision_detection):
//   QTNode        quadtree/qt_node.rs:14-25        -> node_meta, node_bbox, node_nactive
//   QTHazardVec   quadtree/qt_hazard_vec.rs:13-18  -> entries of a node are contiguous, Entire first
//   QTHazard      quadtree/qt_hazard.rs:13-33      -> entry_meta (+ entry_ffbbox for Partial)
//   QTHazPartial  quadtree/qt_partial_hazard.rs:7-12 -> entry_ffbbox + a range of `edges`
//   Hazard        hazards/hazard.rs:16-33          -> haz_meta, haz_bbox, haz_poi, haz_verts (by slot)
//   HazKey        hazards/hazard.rs:8-12 (slotmap) -> slot index = position in the haz_* sections, version in haz_version
#pragma once
#include <stdint.h>

#define JG_LAYOUT_MAGIC 0x4A47424Cu /* "JGBL" */
#define JG_LAYOUT_VERSION 2u
/* flags: the producer evaluated the grid premise of csrc/detect_grid.cuh on the host (GRID_CHECKED) and it holds
 * (GRID_OK): depth <= 5, every node box equals the nested-midpoint grid bit for bit, leaf records fit. */
#define JG_LAYOUT_FLAG_GRID_CHECKED 1u
#define JG_LAYOUT_FLAG_GRID_OK 2u

typedef struct JgLayoutHeader {
    uint32_t magic, version;
    uint32_t total_bytes;
    uint32_t n_nodes;      // node 0 is the root; the 4 children of a node are contiguous (NE,NW,SW,SE)
    uint32_t n_entries;
    uint32_t n_edges;
    uint32_t n_slots;      // hazard slot capacity (SlotMap slots); dead slots have vert_count == 0
    uint32_t n_haz_verts;
    uint32_t root_level;   // CDEConfig::quadtree_depth
    uint32_t cd_threshold;
    uint32_t exterior_slot;
    uint32_t flags;        // JG_LAYOUT_FLAG_*
    float cde_bbox[4];     // = node_bbox[0]
    // byte offsets from the start of the blob
    uint32_t off_node_meta;    // int4  {child_base | -1, entry_begin, n_entire | n_partial << 16, parent | -1}
    uint32_t off_node_bbox;    // float4 {x_min, y_min, x_max, y_max}
    uint32_t off_node_nactive; // u32   QTHazardVec::n_active_edges
    uint32_t off_entry_meta;   // int4  {hazard slot, edge_begin, edge_count, presence (1 Partial, 2 Entire)}
    uint32_t off_entry_ffbbox; // float4 QTHazPartial::ff_bbox (zeros for Entire)
    uint32_t off_edges;        // float4 {sx, sy, ex, ey}
    uint32_t off_haz_meta;     // int4  {scope (0 Exterior, 1 Interior), vert_begin, vert_count, kind}
    uint32_t off_haz_bbox;     // float4 shape bbox
    uint32_t off_haz_poi;      // float4 {poi.x, poi.y, 0, 0}
    uint32_t off_haz_verts;    // float2
    uint32_t off_haz_version;  // u32   SlotMap version of the slot (odd = occupied): what a HazKey must match (jg_filter_upload_keys)
    uint32_t reserved1;
} JgLayoutHeader;
