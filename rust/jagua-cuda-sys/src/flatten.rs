//! `CDEngine` -> the flattened layout blob of `jagua-rs_b200/csrc/layout_soa.h` (version 2): what `HostCDE::flatten`
//! (csrc/cde_host.cpp) does for the repository's own host mirror, for the REAL Rust structures
//! (`quadtree/qt_node.rs:14-25`, `qt_hazard_vec.rs:13-18`, `qt_hazard.rs:13-33`, `qt_partial_hazard.rs:7-12`,
//! `hazards/hazard.rs:16-33`). NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Rust toolchain).
//!
//! Layout rules the kernels rely on (checked again by `jg_layout_upload`):
//!   * nodes are numbered breadth-first: root = 0, the four children of a node are contiguous in quadrant order
//!     (NE, NW, SW, SE = `Rect::quadrants`, rect.rs:141-161), so every child block starts at an index `1 + 4k`;
//!   * the entries of a node keep the order of its `QTHazardVec` (Entire first, then Partial);
//!   * hazard identity on the device = SlotMap SLOT INDEX (`key.data().as_ffi() as u32`); the slot's version travels in
//!     the `haz_version` section so that versioned filters (`jg_filter_upload_keys`) never match a reused slot.
use jagua_rs::collision_detection::hazards::{HazKey, HazardEntity};
use jagua_rs::collision_detection::quadtree::{QTHazPresence, QTNode};
use jagua_rs::collision_detection::CDEngine;
use jagua_rs::geometry::geo_enums::GeoPosition;
use slotmap::Key;

const MAGIC: u32 = 0x4A47_424C; // "JGBL"
const VERSION: u32 = 2;
const HEADER_WORDS: usize = 28; // 12 u32 + 4 f32 + 12 u32 (JgLayoutHeader)

fn slot_of(k: HazKey) -> u32 {
    k.data().as_ffi() as u32
}
fn version_of(k: HazKey) -> u32 {
    (k.data().as_ffi() >> 32) as u32
}
fn n_slots(cde: &CDEngine) -> usize {
    cde.hazards_map.keys().map(|k| slot_of(k) as usize + 1).max().unwrap_or(0)
}
/// u32 words a collector / filter bitset needs for this CDE
pub fn collector_words(cde: &CDEngine) -> usize {
    n_slots(cde).div_ceil(32).max(1)
}
/// slot index -> live HazKey (None for free slots)
pub fn keys_by_slot(cde: &CDEngine) -> Vec<Option<HazKey>> {
    let mut v = vec![None; n_slots(cde)];
    for k in cde.hazards_map.keys() {
        v[slot_of(k) as usize] = Some(k);
    }
    v
}
fn kind_of(e: &HazardEntity) -> i32 {
    match e {
        HazardEntity::PlacedItem { .. } => 0,
        HazardEntity::Exterior => 1,
        HazardEntity::Hole { .. } => 2,
        HazardEntity::InferiorQualityZone { .. } => 3,
    }
}

struct Sections {
    node_meta: Vec<i32>,   // {child_base | -1, entry_begin, n_entire | n_partial << 16, parent | -1}
    node_bbox: Vec<f32>,
    node_nactive: Vec<u32>,
    entry_meta: Vec<i32>,  // {slot, edge_begin, edge_count, presence (1 Partial, 2 Entire)}
    entry_ff: Vec<f32>,
    edges: Vec<f32>,
}

pub fn flatten_cde(cde: &CDEngine) -> Vec<u8> {
    // breadth-first order
    let mut order: Vec<(&QTNode, i32)> = vec![(&cde.quadtree, -1)];
    let mut child_base: Vec<i32> = Vec::new();
    let mut f = 0;
    while f < order.len() {
        let (node, _) = order[f];
        match node.children.as_ref() {
            Some(ch) => {
                child_base.push(order.len() as i32);
                for c in ch.iter() {
                    order.push((c, f as i32));
                }
            }
            None => child_base.push(-1),
        }
        f += 1;
    }
    let mut s = Sections { node_meta: vec![], node_bbox: vec![], node_nactive: vec![], entry_meta: vec![], entry_ff: vec![], edges: vec![] };
    for (f, (node, parent)) in order.iter().enumerate() {
        let entry_begin = (s.entry_meta.len() / 4) as i32;
        let (mut n_entire, mut n_partial) = (0i32, 0i32);
        for hz in node.hazards.iter() {
            let edge_begin = (s.edges.len() / 4) as i32;
            match &hz.presence {
                QTHazPresence::None => continue, // never stored (qt_hazard_vec.rs:25)
                QTHazPresence::Entire => {
                    n_entire += 1;
                    s.entry_meta.extend([slot_of(hz.hkey) as i32, edge_begin, 0, 2]);
                    s.entry_ff.extend([0.0; 4]);
                }
                QTHazPresence::Partial(p) => {
                    n_partial += 1;
                    for e in p.edges.iter() {
                        s.edges.extend([e.start.0, e.start.1, e.end.0, e.end.1]);
                    }
                    s.entry_meta.extend([slot_of(hz.hkey) as i32, edge_begin, p.edges.len() as i32, 1]);
                    s.entry_ff.extend([p.ff_bbox.x_min, p.ff_bbox.y_min, p.ff_bbox.x_max, p.ff_bbox.y_max]);
                }
            }
        }
        s.node_meta.extend([child_base[f], entry_begin, n_entire | (n_partial << 16), *parent]);
        s.node_bbox.extend([node.bbox.x_min, node.bbox.y_min, node.bbox.x_max, node.bbox.y_max]);
        s.node_nactive.push(node.hazards.n_active_edges() as u32);
    }
    // hazards by slot
    let ns = n_slots(cde);
    let mut haz_meta = vec![0i32; 4 * ns]; // {scope (0 Exterior, 1 Interior), vert_begin, vert_count, kind}; dead slot: count 0
    let mut haz_bbox = vec![0f32; 4 * ns];
    let mut haz_poi = vec![0f32; 4 * ns];
    let mut haz_version = vec![0u32; ns];
    let mut haz_verts: Vec<f32> = Vec::new();
    let mut exterior_slot = 0u32;
    for (k, h) in cde.hazards_map.iter() {
        let i = slot_of(k) as usize;
        haz_version[i] = version_of(k);
        let scope = if h.entity.scope() == GeoPosition::Exterior { 0 } else { 1 };
        haz_meta[4 * i..4 * i + 4].copy_from_slice(&[scope, (haz_verts.len() / 2) as i32, h.shape.vertices.len() as i32, kind_of(&h.entity)]);
        haz_bbox[4 * i..4 * i + 4].copy_from_slice(&[h.shape.bbox.x_min, h.shape.bbox.y_min, h.shape.bbox.x_max, h.shape.bbox.y_max]);
        haz_poi[4 * i..4 * i + 2].copy_from_slice(&[h.shape.poi.center.0, h.shape.poi.center.1]);
        for p in h.shape.vertices.iter() {
            haz_verts.extend([p.0, p.1]);
        }
        if matches!(h.entity, HazardEntity::Exterior) && exterior_slot == 0 {
            exterior_slot = i as u32; // first Exterior in slot order (cd_engine.rs:44-48)
        }
    }
    // assemble: header, then 16-byte aligned sections
    let mut blob: Vec<u8> = vec![0u8; HEADER_WORDS * 4];
    let mut offs = [0u32; 11];
    let mut put = |blob: &mut Vec<u8>, bytes: &[u8], slot: usize, offs: &mut [u32; 11]| {
        while blob.len() % 16 != 0 {
            blob.push(0);
        }
        offs[slot] = blob.len() as u32;
        blob.extend_from_slice(bytes);
    };
    fn bytes_of<T: Copy>(v: &[T]) -> &[u8] {
        unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u8, std::mem::size_of_val(v)) }
    }
    put(&mut blob, bytes_of(&s.node_meta), 0, &mut offs);
    put(&mut blob, bytes_of(&s.node_bbox), 1, &mut offs);
    put(&mut blob, bytes_of(&s.node_nactive), 2, &mut offs);
    put(&mut blob, bytes_of(&s.entry_meta), 3, &mut offs);
    put(&mut blob, bytes_of(&s.entry_ff), 4, &mut offs);
    put(&mut blob, bytes_of(&s.edges), 5, &mut offs);
    put(&mut blob, bytes_of(&haz_meta), 6, &mut offs);
    put(&mut blob, bytes_of(&haz_bbox), 7, &mut offs);
    put(&mut blob, bytes_of(&haz_poi), 8, &mut offs);
    put(&mut blob, bytes_of(&haz_verts), 9, &mut offs);
    put(&mut blob, bytes_of(&haz_version), 10, &mut offs);
    while blob.len() % 16 != 0 {
        blob.push(0);
    }
    let root = &cde.quadtree;
    let mut hdr: Vec<u32> = vec![
        MAGIC, VERSION, blob.len() as u32, order.len() as u32, (s.entry_meta.len() / 4) as u32, (s.edges.len() / 4) as u32, ns as u32,
        (haz_verts.len() / 2) as u32, cde.config.quadtree_depth as u32, cde.config.cd_threshold as u32, exterior_slot,
        0, // flags: the engine evaluates the grid premise itself when JG_LAYOUT_FLAG_GRID_CHECKED is absent
        root.bbox.x_min.to_bits(), root.bbox.y_min.to_bits(), root.bbox.x_max.to_bits(), root.bbox.y_max.to_bits(),
    ];
    hdr.extend(offs);
    hdr.push(0); // reserved1
    debug_assert_eq!(hdr.len(), HEADER_WORDS);
    blob[..HEADER_WORDS * 4].copy_from_slice(bytes_of(&hdr));
    blob
}
