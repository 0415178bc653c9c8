// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see geom.hpp header).
// extern "C" surface over the oracle for ctypes (tests/, smoke(), bench.py's CPU legs).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "entities.hpp"

namespace orc {
#ifdef ORC_COUNT
thread_local uint64_t g_flops = 0;
#endif

struct World {
    CDEConfig cfg;
    std::vector<Item> items;
    std::unique_ptr<Layout> layout;
    float strip_width = 0, strip_height = 0;
    bool is_strip = false;
    // container under construction
    std::vector<Point> pend_outer;
    std::vector<QualityZoneShape> pend_zones;
    std::vector<SPolygon> pend_holes;
    Rect pend_bbox{0, 0, 0, 0};
    std::vector<Hazard> pend_static;
    std::vector<std::unique_ptr<CDESnapshot>> snapshots;
    std::vector<std::vector<std::pair<int, DTransformation>>> snapshot_items;
    std::string last_error;
};

static std::vector<Point> to_points(const float* xy, int n) {
    std::vector<Point> p(n);
    for (int i = 0; i < n; i++) p[i] = {xy[2 * i], xy[2 * i + 1]};
    return p;
}
}  // namespace orc

using namespace orc;

#define ORC_TRY(w) try {
#define ORC_CATCH(w, ret)                   \
    }                                       \
    catch (const std::exception& e) {       \
        (w)->last_error = e.what();         \
        return ret;                         \
    }

extern "C" {

World* orc_world_new(int quadtree_depth, int cd_threshold, int n_ff_poles, int n_ff_piers) {
    World* w = new World();
    w->cfg.quadtree_depth = (uint8_t)quadtree_depth;
    w->cfg.cd_threshold = (uint8_t)cd_threshold;
    w->cfg.item_surrogate_config.n_ff_poles = n_ff_poles;
    w->cfg.item_surrogate_config.n_ff_piers = n_ff_piers;
    return w;
}
void orc_world_free(World* w) { delete w; }
const char* orc_last_error(World* w) { return w->last_error.c_str(); }
int orc_counting_enabled() {
#ifdef ORC_COUNT
    return 1;
#else
    return 0;
#endif
}

// ---- items -------------------------------------------------------------------------------
// n_orient < 0: continuous rotation (allowed_orientations absent); orient in degrees.
int orc_add_item(World* w, const float* xy, int n, const float* orient_deg, int n_orient, int min_quality,
                 int demand) {
    ORC_TRY(w)
    int id = (int)w->items.size();
    w->items.push_back(import_item(id, to_points(xy, n), orient_deg, n_orient, min_quality, demand,
                                   w->cfg.item_surrogate_config));
    return id;
    ORC_CATCH(w, -1)
}
int orc_n_items(World* w) { return (int)w->items.size(); }
int orc_item_n_vertices(World* w, int id) { return w->items[id].shape_cd.n_vertices(); }
int orc_item_n_poles(World* w, int id) { return (int)w->items[id].shape_cd.surrogate.poles.size(); }
int orc_item_n_piers(World* w, int id) { return (int)w->items[id].shape_cd.surrogate.piers.size(); }
int orc_item_n_rotations(World* w, int id) {
    const Item& it = w->items[id];
    return it.rot_kind == RotKind::Continuous ? -1 : (it.rot_kind == RotKind::None ? 0 : (int)it.rotations.size());
}
// info: [poi_x, poi_y, poi_r, area, diameter, bbox(4), pre_tx, pre_ty, min_quality, demand] = 13 floats
void orc_item_get(World* w, int id, float* verts_xy, float* poles_xyr, float* piers, float* rotations,
                  float* info) {
    const Item& it = w->items[id];
    const SPolygon& s = it.shape_cd;
    for (int i = 0; i < s.n_vertices(); i++) {
        verts_xy[2 * i] = s.vertices[i].x;
        verts_xy[2 * i + 1] = s.vertices[i].y;
    }
    for (size_t i = 0; i < s.surrogate.poles.size(); i++) {
        poles_xyr[3 * i] = s.surrogate.poles[i].center.x;
        poles_xyr[3 * i + 1] = s.surrogate.poles[i].center.y;
        poles_xyr[3 * i + 2] = s.surrogate.poles[i].radius;
    }
    for (size_t i = 0; i < s.surrogate.piers.size(); i++) {
        piers[4 * i] = s.surrogate.piers[i].start.x;
        piers[4 * i + 1] = s.surrogate.piers[i].start.y;
        piers[4 * i + 2] = s.surrogate.piers[i].end.x;
        piers[4 * i + 3] = s.surrogate.piers[i].end.y;
    }
    for (size_t i = 0; i < it.rotations.size(); i++) rotations[i] = it.rotations[i];
    float v[13] = {s.poi.center.x, s.poi.center.y, s.poi.radius, s.area, s.diameter, s.bbox.x_min,
                   s.bbox.y_min, s.bbox.x_max, s.bbox.y_max, it.pre_transform.tx, it.pre_transform.ty,
                   (float)it.min_quality, (float)it.demand};
    std::memcpy(info, v, sizeof(v));
}
// Inject piers (tests only; the reference generates them in fail_fast/piers.rs, out of scope)
void orc_item_set_piers(World* w, int id, const float* piers, int n) {
    auto& v = w->items[id].shape_cd.surrogate.piers;
    v.clear();
    for (int i = 0; i < n; i++) v.push_back({{piers[4 * i], piers[4 * i + 1]}, {piers[4 * i + 2], piers[4 * i + 3]}});
}
// Replace an item by explicit internal data (used to load workload fixtures without re-importing)
int orc_add_item_raw(World* w, const float* verts_xy, int n, const float* poi_xyr, const float* poles_xyr,
                     int n_poles, const float* piers, int n_piers, const float* rotations, int n_rot,
                     int min_quality) {
    ORC_TRY(w)
    Item it;
    it.id = (int)w->items.size();
    SPolygon& s = it.shape_cd;
    s.vertices = to_points(verts_xy, n);
    s.bbox = SPolygon::generate_bounding_box(s.vertices);
    s.area = SPolygon::calculate_area(s.vertices);
    s.diameter = calculate_diameter(s.vertices);
    s.poi = {{poi_xyr[0], poi_xyr[1]}, poi_xyr[2]};
    s.has_surrogate = true;
    s.surrogate.config = w->cfg.item_surrogate_config;
    for (int i = 0; i < n_poles; i++)
        s.surrogate.poles.push_back({{poles_xyr[3 * i], poles_xyr[3 * i + 1]}, poles_xyr[3 * i + 2]});
    for (int i = 0; i < n_piers; i++)
        s.surrogate.piers.push_back({{piers[4 * i], piers[4 * i + 1]}, {piers[4 * i + 2], piers[4 * i + 3]}});
    it.shape_orig = s;
    if (n_rot < 0) it.rot_kind = RotKind::Continuous;
    else if (n_rot == 0) it.rot_kind = RotKind::None;
    else {
        it.rot_kind = RotKind::Discrete;
        it.rotations.assign(rotations, rotations + n_rot);
    }
    it.min_quality = min_quality;
    w->items.push_back(std::move(it));
    return (int)w->items.size() - 1;
    ORC_CATCH(w, -1)
}

// ---- containers / layout -----------------------------------------------------------------
int orc_set_strip(World* w, float width, float height) {
    ORC_TRY(w)
    w->is_strip = true;
    w->strip_width = width;
    w->strip_height = height;
    w->layout = std::make_unique<Layout>(strip_container(width, height, w->cfg));
    return 0;
    ORC_CATCH(w, -1)
}
float orc_strip_width(World* w) { return w->strip_width; }
void orc_container_begin(World* w, const float* outer_xy, int n) {
    w->pend_outer = to_points(outer_xy, n);
    w->pend_zones.clear();
    w->pend_holes.clear();
}
int orc_container_add_zone(World* w, int quality, const float* xy, int n) {
    ORC_TRY(w)
    w->pend_zones.push_back({quality, import_simple_polygon(to_points(xy, n))});
    return 0;
    ORC_CATCH(w, -1)
}
int orc_container_add_hole(World* w, const float* xy, int n) {
    ORC_TRY(w)
    w->pend_holes.push_back(import_simple_polygon(to_points(xy, n)));
    return 0;
    ORC_CATCH(w, -1)
}
int orc_container_finish(World* w) {
    ORC_TRY(w)
    w->is_strip = false;
    w->layout = std::make_unique<Layout>(
        make_container(import_simple_polygon(w->pend_outer), w->pend_zones, w->pend_holes, w->cfg));
    return 0;
    ORC_CATCH(w, -1)
}
// Build a container + CDE directly from raw static hazards (workload fixtures): begin(bbox) /
// add_static (first must be the Exterior) / finish. Shapes are taken as-is (already internal, CCW).
static SPolygon raw_polygon(const float* xy, int n, const float* poi_xyr) {
    SPolygon s;
    s.vertices = to_points(xy, n);
    s.bbox = SPolygon::generate_bounding_box(s.vertices);
    s.area = SPolygon::calculate_area(s.vertices);
    s.diameter = 0.0f;
    s.poi = {{poi_xyr[0], poi_xyr[1]}, poi_xyr[2]};
    return s;
}
void orc_cde_begin(World* w, const float* bbox4) {
    w->pend_bbox = {bbox4[0], bbox4[1], bbox4[2], bbox4[3]};
    w->pend_static.clear();
}
void orc_cde_add_static(World* w, int kind, int64_t id, int64_t quality, const float* xy, int n,
                        const float* poi_xyr) {
    HazardEntity e;
    e.kind = (EntityKind)kind;
    e.id = id;
    e.quality = quality;
    w->pend_static.push_back({e, raw_polygon(xy, n, poi_xyr), false});
}
int orc_cde_finish(World* w) {
    ORC_TRY(w)
    Container c;
    c.outer_cd = w->pend_static.at(0).shape;
    c.base_cde = std::make_shared<CDEngine>(w->pend_bbox, w->pend_static, w->cfg);
    w->is_strip = false;
    w->layout = std::make_unique<Layout>(std::move(c));
    return 0;
    ORC_CATCH(w, -1)
}
// Raw (de)registration of a hazard on the current layout's CDE (cd_engine.rs:59-96): shapes are taken
// as-is. Used to rebuild a layout from a workload fixture with the same slot numbering.
uint64_t orc_register_hazard_raw(World* w, int kind, int64_t id, int64_t quality, uint64_t pk, const float* xy,
                                 int n, const float* poi_xyr, int dynamic) {
    ORC_TRY(w)
    HazardEntity e;
    e.kind = (EntityKind)kind;
    e.id = id;
    e.quality = quality;
    e.pk = pk;
    return w->layout->cde.register_hazard({e, raw_polygon(xy, n, poi_xyr), dynamic != 0}).as_u64();
    ORC_CATCH(w, 0)
}
int orc_deregister_hazard(World* w, uint64_t key) {
    ORC_TRY(w)
    w->layout->cde.deregister_hazard_by_key(HazKey::from_u64(key));
    return 0;
    ORC_CATCH(w, -1)
}
int orc_cde_save(World* w) {
    w->snapshots.push_back(std::make_unique<CDESnapshot>(w->layout->cde.save()));
    return (int)w->snapshots.size() - 1;
}
int orc_cde_restore(World* w, int snap) {
    ORC_TRY(w)
    w->layout->cde.restore(*w->snapshots.at(snap));
    return 0;
    ORC_CATCH(w, -1)
}
void orc_cde_bbox(World* w, float* out4) {
    Rect b = w->layout->cde.bbox();
    out4[0] = b.x_min;
    out4[1] = b.y_min;
    out4[2] = b.x_max;
    out4[3] = b.y_max;
}

uint64_t orc_place_item(World* w, int item_id, float rot, float tx, float ty) {
    ORC_TRY(w)
    return w->layout->place_item(w->items[item_id], {rot, tx, ty});
    ORC_CATCH(w, 0)
}
int orc_remove_item(World* w, uint64_t pk) {
    ORC_TRY(w)
    w->layout->remove_item(pk);
    return 0;
    ORC_CATCH(w, -1)
}
int orc_n_placed(World* w) { return (int)w->layout->n_placed(); }
// out: pk[], item_id[], dt[3*n]
void orc_placed_get(World* w, uint64_t* pks, int* item_ids, float* dts) {
    int i = 0;
    w->layout->for_each_placed([&](uint64_t pk, const PlacedItem& pi) {
        pks[i] = pk;
        item_ids[i] = pi.item_id;
        dts[3 * i] = pi.d_transf.rotation;
        dts[3 * i + 1] = pi.d_transf.tx;
        dts[3 * i + 2] = pi.d_transf.ty;
        i++;
    });
}
int orc_is_feasible(World* w) {
    ORC_TRY(w)
    return w->layout->is_feasible() ? 1 : 0;
    ORC_CATCH(w, -1)
}
int orc_change_strip_width(World* w, float width) {
    ORC_TRY(w)
    w->strip_width = width;
    w->layout->swap_container(strip_container(width, w->strip_height, w->cfg));
    return 0;
    ORC_CATCH(w, -1)
}
// CDE-level save/restore (cd_engine.rs:98-149) + placed-items bookkeeping (layout.rs:52-69)
int orc_save(World* w) {
    w->snapshots.push_back(std::make_unique<CDESnapshot>(w->layout->cde.save()));
    return (int)w->snapshots.size() - 1;
}
int orc_restore(World* w, int snap) {
    ORC_TRY(w)
    // Layout::restore: placed_items = snapshot.placed_items; cde.restore(snapshot)
    const CDESnapshot& s = *w->snapshots[snap];
    // rebuild the placed-items slotmap to hold exactly the snapshot's PlacedItem entities
    Layout& L = *w->layout;
    std::vector<bool> keep(L.placed.size(), false);
    for (const Hazard& h : s.dynamic_hazards) {
        uint32_t idx = (uint32_t)h.entity.pk;
        if (idx >= L.placed.size()) L.placed.resize(idx + 1);
        if (idx >= keep.size()) keep.resize(idx + 1, false);
        keep[idx] = true;
        if (!L.placed[idx].pi || L.placed[idx].version != (uint32_t)(h.entity.pk >> 32)) {
            auto pi = std::make_unique<PlacedItem>();
            pi->item_id = (int)h.entity.id;
            pi->d_transf = h.entity.dt;
            pi->shape = h.shape;
            L.placed[idx].pi = std::move(pi);
            L.placed[idx].version = (uint32_t)(h.entity.pk >> 32);
        }
    }
    for (size_t i = 0; i < L.placed.size(); i++)
        if (!keep[i] && L.placed[i].pi) {
            L.placed[i].pi.reset();
            L.placed[i].version += 1;
        }
    // rebuild the free list (LIFO over free slots, highest index first so low indices pop first)
    L.free_head = UINT32_MAX;
    for (size_t i = L.placed.size(); i-- > 0;)
        if (!L.placed[i].pi) {
            if (L.placed[i].version % 2 == 1) L.placed[i].version += 1;
            L.placed[i].next_free = L.free_head;
            L.free_head = (uint32_t)i;
        }
    L.cde.restore(s);
    return 0;
    ORC_CATCH(w, -1)
}

// ---- hazards dump ------------------------------------------------------------------------
int orc_n_hazards(World* w) { return (int)w->layout->cde.hazards_map.n; }
int orc_haz_capacity(World* w) { return (int)w->layout->cde.hazards_map.capacity(); }
// per hazard: keys[u64], meta[i64 x 6] = kind, id, quality, scope(0 ext,1 int), dynamic, n_verts;
// fl[f32 x 10] = poi x,y,r, bbox(4), dt(3); pk[u64]
void orc_hazards_get(World* w, uint64_t* keys, int64_t* meta, float* fl, uint64_t* pks) {
    int i = 0;
    w->layout->cde.hazards_map.for_each([&](HazKey k, const Hazard& h) {
        keys[i] = k.as_u64();
        meta[6 * i + 0] = (int64_t)h.entity.kind;
        meta[6 * i + 1] = h.entity.id;
        meta[6 * i + 2] = h.entity.quality;
        meta[6 * i + 3] = h.entity.scope() == GeoPosition::Interior ? 1 : 0;
        meta[6 * i + 4] = h.dynamic ? 1 : 0;
        meta[6 * i + 5] = h.shape.n_vertices();
        float v[10] = {h.shape.poi.center.x, h.shape.poi.center.y, h.shape.poi.radius, h.shape.bbox.x_min,
                       h.shape.bbox.y_min, h.shape.bbox.x_max, h.shape.bbox.y_max, h.entity.dt.rotation,
                       h.entity.dt.tx, h.entity.dt.ty};
        std::memcpy(fl + 10 * i, v, sizeof(v));
        pks[i] = h.entity.pk;
        i++;
    });
}
int orc_hazard_vertices(World* w, uint64_t key, float* xy) {
    ORC_TRY(w)
    const Hazard& h = w->layout->cde.hazards_map.at(HazKey::from_u64(key));
    for (int i = 0; i < h.shape.n_vertices(); i++) {
        xy[2 * i] = h.shape.vertices[i].x;
        xy[2 * i + 1] = h.shape.vertices[i].y;
    }
    return h.shape.n_vertices();
    ORC_CATCH(w, -1)
}
uint64_t orc_hkey_exterior(World* w) { return w->layout->cde.hkey_exterior.as_u64(); }
// filter.rs:21-41
int orc_filter_qzones(World* w, int required_quality, uint32_t* bits, int words) {
    HazardFilter f = filter_from_irrelevant_qzones(required_quality, w->layout->cde.hazards_map);
    for (int i = 0; i < words; i++) bits[i] = i < (int)f.bits.size() ? f.bits[i] : 0;
    return 0;
}

// ---- canonical quadtree dump (compared byte-for-byte with the product's own dump) ---------
static void dump_node(const QTNode& n, std::vector<uint32_t>& out) {
    auto pf = [&](float f) {
        uint32_t u;
        std::memcpy(&u, &f, 4);
        out.push_back(u);
    };
    out.push_back(n.level);
    pf(n.bbox.x_min), pf(n.bbox.y_min), pf(n.bbox.x_max), pf(n.bbox.y_max);
    out.push_back(n.children ? 1 : 0);
    out.push_back((uint32_t)n.hazards.hazards.size());
    out.push_back((uint32_t)n.hazards.n_active_edges);
    std::vector<const QTHazard*> es;
    for (auto& h : n.hazards.hazards) es.push_back(&h);
    std::stable_sort(es.begin(), es.end(), [](const QTHazard* a, const QTHazard* b) {
        if (a->presence != b->presence) return (int)a->presence > (int)b->presence;
        return a->hkey.idx < b->hkey.idx;
    });
    for (const QTHazard* h : es) {
        out.push_back(h->hkey.idx);
        out.push_back((uint32_t)h->presence);
        if (h->presence == Presence::Partial) {
            pf(h->partial.ff_bbox.x_min), pf(h->partial.ff_bbox.y_min), pf(h->partial.ff_bbox.x_max),
                pf(h->partial.ff_bbox.y_max);
            out.push_back((uint32_t)h->partial.edges.size());
            for (const Edge& e : h->partial.edges) pf(e.start.x), pf(e.start.y), pf(e.end.x), pf(e.end.y);
        } else {
            for (int i = 0; i < 4; i++) out.push_back(0);
            out.push_back(0);
        }
    }
    if (n.children)
        for (const QTNode& c : *n.children) dump_node(c, out);
}
// returns the number of u32 words; writes at most cap words
int64_t orc_tree_dump(World* w, uint32_t* out, int64_t cap) {
    std::vector<uint32_t> v;
    dump_node(w->layout->cde.quadtree, v);
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
}
// Reference debug invariants (util/assertions.rs): returns 0 when all hold, else a code.
static int check_node(const QTNode& n, const QTNode* parent) {
    // qt_hazard_vec.rs:102-120: sorted Entire -> Partial, n_active_edges exact
    size_t edges = 0;
    int prev = 3;
    for (auto& h : n.hazards.hazards) {
        if ((int)h.presence > prev) return 1;
        if (h.presence == Presence::None) return 2;
        prev = (int)h.presence;
        edges += h.n_edges();
    }
    if (edges != n.hazards.n_active_edges) return 3;
    // assertions.rs:52-89: child hazards are a subset of the parent's
    if (parent)
        for (auto& h : n.hazards.hazards) {
            bool found = false;
            for (auto& ph : parent->hazards.hazards) found = found || ph.hkey == h.hkey;
            if (!found) return 4;
        }
    // qt_node.rs:60-75,87-89: children iff at least one Partial hazard and level > 0
    bool should = n.level > 0 && !n.hazards.no_partial_hazards();
    if (should != (bool)n.children) return 5;
    if (n.children) {
        Rect q[4];
        n.bbox.quadrants(q);
        for (int i = 0; i < 4; i++) {
            if (!((*n.children)[i].bbox == q[i])) return 6;
            if ((*n.children)[i].level != n.level - 1) return 7;
            int r = check_node((*n.children)[i], &n);
            if (r) return r;
        }
    }
    return 0;
}
int orc_check_invariants(World* w) { return check_node(w->layout->cde.quadtree, nullptr); }
// assertions.rs:91-104 layout_qt_matches_fresh_qt: rebuild from scratch and compare dumps
int orc_matches_fresh(World* w) {
    ORC_TRY(w)
    const Layout& L = *w->layout;
    CDEngine fresh(*L.container.base_cde);
    // re-register dynamic hazards in slot order so that keys match slot-for-slot
    std::vector<std::pair<HazKey, Hazard>> dyn;
    L.cde.hazards_map.for_each([&](HazKey k, const Hazard& h) {
        if (h.dynamic) dyn.push_back({k, h});
    });
    // fresh registration hands out new slots; compare per node sets of (entity, presence-kind)
    std::vector<std::pair<HazKey, HazKey>> keymap;  // (orig, fresh)
    for (auto& kv : dyn) keymap.push_back({kv.first, fresh.register_hazard(kv.second)});
    std::function<int(const QTNode&, const QTNode&)> cmp = [&](const QTNode& a, const QTNode& b) -> int {
        if (a.hazards.hazards.size() != b.hazards.hazards.size()) return 1;
        if ((bool)a.children != (bool)b.children) return 2;
        for (auto& ha : a.hazards.hazards) {
            bool found = false;
            for (auto& hb : b.hazards.hazards)
                if (ha.entity == hb.entity && ha.presence == hb.presence) {
                    found = true;
                    if (ha.presence == Presence::Partial) {
                        if (!(ha.partial.edges == hb.partial.edges)) return 3;
                        if (!(ha.partial.ff_bbox == hb.partial.ff_bbox)) return 4;
                    }
                }
            if (!found) return 5;
        }
        if (a.children)
            for (int i = 0; i < 4; i++) {
                int r = cmp((*a.children)[i], (*b.children)[i]);
                if (r) return r;
            }
        return 0;
    };
    return cmp(L.cde.quadtree, fresh.quadtree);
    ORC_CATCH(w, -1)
}

// ---- LBF fillers -------------------------------------------------------------------------
// lbf_spp.rs:37-123. Returns the number of placed items, -1 on error.
int orc_lbf_solve_spp(World* w, int n_samples, float ls_frac, uint64_t seed, int max_items) {
    ORC_TRY(w)
    Rng rng(seed);
    LBFParams prm{n_samples, ls_frac};
    std::vector<int> demand;
    for (auto& it : w->items) demand.push_back(it.demand);
    int placed = 0;
    for (int item_id : item_placement_order(w->items)) {
        const Item& item = w->items[item_id];
        int guard = 0;
        while (demand[item_id] > 0 && placed < max_items) {
            const CDEngine& cde = w->layout->cde;
            HazardFilter f = item_filter(cde, item);
            DTransformation dt;
            if (lbf_search(cde, item, prm, rng, f, dt)) {
                w->layout->place_item(item, dt);
                demand[item_id]--;
                placed++;
            } else {
                if (++guard > 100) throw std::runtime_error("strip width is running away");
                w->strip_width *= 1.1f;
                w->layout->swap_container(strip_container(w->strip_width, w->strip_height, w->cfg));
            }
        }
    }
    // problem.rs:39-57 fit_strip
    float x_max = -INFINITY;
    w->layout->for_each_placed([&](uint64_t, const PlacedItem& pi) { x_max = std::fmax(x_max, pi.shape.bbox.x_max); });
    if (placed > 0) {
        w->strip_width = x_max * 1.00001f;
        w->layout->swap_container(strip_container(w->strip_width, w->strip_height, w->cfg));
    }
    return placed;
    ORC_CATCH(w, -1)
}
// lbf_bpp.rs:39-74 restricted to the single open layout of this world (one bin)
int orc_lbf_fill_bin(World* w, int n_samples, float ls_frac, uint64_t seed, int max_items) {
    ORC_TRY(w)
    Rng rng(seed);
    LBFParams prm{n_samples, ls_frac};
    int placed = 0;
    for (int item_id : item_placement_order(w->items)) {
        const Item& item = w->items[item_id];
        int demand = item.demand;
        while (demand > 0 && placed < max_items) {
            const CDEngine& cde = w->layout->cde;
            HazardFilter f = item_filter(cde, item);
            DTransformation dt;
            if (!lbf_search(cde, item, prm, rng, f, dt)) break;
            w->layout->place_item(item, dt);
            demand--;
            placed++;
        }
    }
    return placed;
    ORC_CATCH(w, -1)
}

// ---- batched queries (the CPU baseline loop; lbf/benches/ci_bench.rs:56-70,107-121) -------
// The reference is single-threaded on the query path; "rayon-parallel over the same queries" (BASELINE north_star) is a
// parallel-for over independent queries. Here: a persistent pool of worker threads (created once, like rayon's global
// pool) that take chunks of ORC_CHUNK queries from an atomic counter (dynamic balancing: per-query cost varies a lot).
#define ORC_CHUNK 2048
namespace {
struct ThreadPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    const std::function<void(int)>* job = nullptr;
    std::atomic<uint64_t> gen{0};
    std::atomic<int> want{0};
    int running = 0;
    bool stop = false;
    void worker(int k) {
        uint64_t seen = 0;
        for (;;) {
            // back-to-back parallel regions (the benchmark loop) find the workers still spinning: a blocked thread can take
            // tens of milliseconds to wake up in a container, longer than a small batch lasts
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            while (std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(20)) {
                if (gen.load(std::memory_order_acquire) != seen && k < want.load(std::memory_order_acquire)) {
                    got = true;
                    break;
                }
                std::this_thread::yield();
            }
            const std::function<void(int)>* j;
            {
                std::unique_lock<std::mutex> lk(m);
                if (!got) cv_go.wait(lk, [&] { return stop || (gen.load() != seen && k < want.load()); });
                if (stop) return;
                if (gen.load() == seen || k >= want.load()) continue;  // (the region ended while this thread was still spinning)
                seen = gen.load();
                j = job;
            }
            (*j)(k);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--running == 0) cv_done.notify_all();
            }
        }
    }
    // runs fn(k) for k in [0, n) on n threads (the caller is thread n - 1) and waits
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 1) {
            fn(0);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m);
            while ((int)th.size() < n - 1) {
                const int k = (int)th.size();
                th.emplace_back([this, k] { worker(k); });
            }
            job = &fn;
            running = n - 1;
            want.store(n - 1, std::memory_order_release);
            gen.fetch_add(1, std::memory_order_release);
        }
        cv_go.notify_all();
        fn(n - 1);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return running == 0; });
        want.store(0);
    }
    ~ThreadPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& t : th) t.join();
    }
};
ThreadPool& pool() {
    static ThreadPool* p = new ThreadPool();  // leaked on purpose: no destructor race at process exit
    return *p;
}
// segment of query q (segments are contiguous, offsets non-decreasing)
inline int seg_of(const int64_t* offs, int n_seg, int64_t q, int hint) {
    int k = hint;
    while (k + 1 < n_seg && q >= offs[k + 1]) k++;
    while (k > 0 && q < offs[k]) k--;
    return k;
}
}  // namespace

// transforms: n x (cos, sin, tx, ty). mode bit0: poly only (surrogate disabled, quadtree_bench.rs:73-124)
// out byte: bit0 surrogate hit, bit1 poly hit (only evaluated when bit0 == 0), bit7 reference panics
// Multi-item batch: segment k = queries seg_offsets[k] .. seg_offsets[k+1] of item item_ids[k].
// Returns algorithmic flops (ORC_COUNT builds) or 0.
uint64_t orc_detect_multi(World* w, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* t4,
                          const uint32_t* filter_bits, int filter_words, uint8_t* out, int nthreads, int mode) {
    const CDEngine& cde = w->layout->cde;
    HazardFilter filter;
    if (filter_bits) filter.bits.assign(filter_bits, filter_bits + filter_words);
    const int64_t n = seg_offsets[n_seg];
    std::atomic<uint64_t> flops{0};
    std::atomic<int64_t> next{0};
    auto work = [&](int) {
#ifdef ORC_COUNT
        g_flops = 0;
#endif
        int k = 0, buf_item = -1;
        SPolygon buffer;
        for (;;) {
            const int64_t lo = next.fetch_add(ORC_CHUNK);
            if (lo >= n) break;
            const int64_t hi = std::min<int64_t>(n, lo + ORC_CHUNK);
            for (int64_t i = lo; i < hi; i++) {
                k = seg_of(seg_offsets, n_seg, i, k);
                const Item& item = w->items[item_ids[k]];
                if (buf_item != item_ids[k]) {
                    buffer = item.shape_cd;
                    buffer.has_surrogate = false;
                    buf_item = item_ids[k];
                }
                Transformation t = from_cstt(t4[4 * i], t4[4 * i + 1], t4[4 * i + 2], t4[4 * i + 3]);
                uint8_t v = 0;
                if (!(mode & 1) && cde.detect_surrogate_collision(item.shape_cd.surrogate, t, filter)) v |= 1;
                if (!(v & 1)) {
                    buffer.transform_from(item.shape_cd, t);
                    try {
                        if (cde.detect_poly_collision(buffer, filter)) v |= 2;
                    } catch (const PanicInReference&) {
                        v |= 0x80;
                    }
                }
                out[i] = v;
            }
        }
#ifdef ORC_COUNT
        flops += g_flops;
#endif
    };
    pool().run(nthreads < 1 ? 1 : nthreads, work);
    return flops.load();
}
uint64_t orc_detect_batch(World* w, int item_id, const float* t4, int64_t n, const uint32_t* filter_bits,
                          int filter_words, uint8_t* out, int nthreads, int mode) {
    const int64_t offs[2] = {0, n};
    return orc_detect_multi(w, 1, &item_id, offs, t4, filter_bits, filter_words, out, nthreads, mode);
}
// out_bits: n x words_per_query u32 (bit = hazard slot index). mode bit0: poly only. seed_bits (or null): the collectors
// start from these sets instead of empty ones (a collector passed in non-empty: its hazards are skipped by
// `contains_key`, qt_node.rs:172 / cd_engine.rs:279, and stay in the result).
uint64_t orc_collect_multi(World* w, int n_seg, const int* item_ids, const int64_t* seg_offsets, const float* t4,
                           const uint32_t* seed_bits, uint32_t* out_bits, int words_per_query, int nthreads, int mode) {
    const CDEngine& cde = w->layout->cde;
    const int64_t n = seg_offsets[n_seg];
    std::atomic<uint64_t> flops{0};
    std::atomic<int64_t> next{0};
    auto work = [&](int) {
        HazardCollector col(std::max<size_t>(cde.hazards_map.capacity(), (size_t)words_per_query * 32));
#ifdef ORC_COUNT
        g_flops = 0;
#endif
        int k = 0, buf_item = -1;
        SPolygon buffer;
        for (;;) {
            const int64_t lo = next.fetch_add(ORC_CHUNK);
            if (lo >= n) break;
            const int64_t hi = std::min<int64_t>(n, lo + ORC_CHUNK);
            for (int64_t i = lo; i < hi; i++) {
                k = seg_of(seg_offsets, n_seg, i, k);
                const Item& item = w->items[item_ids[k]];
                if (buf_item != item_ids[k]) {
                    buffer = item.shape_cd;
                    buffer.has_surrogate = false;
                    buf_item = item_ids[k];
                }
                Transformation t = from_cstt(t4[4 * i], t4[4 * i + 1], t4[4 * i + 2], t4[4 * i + 3]);
                col.clear();
                if (seed_bits)
                    for (int j = 0; j < words_per_query; j++) col.bits[j] = seed_bits[i * words_per_query + j];
                if (!(mode & 1)) cde.collect_surrogate_collisions(item.shape_cd.surrogate, t, col);
                buffer.transform_from(item.shape_cd, t);
                cde.collect_poly_collisions(buffer, col);
                for (int j = 0; j < words_per_query; j++) out_bits[i * words_per_query + j] = col.bits[j];
            }
        }
#ifdef ORC_COUNT
        flops += g_flops;
#endif
    };
    pool().run(nthreads < 1 ? 1 : nthreads, work);
    return flops.load();
}
uint64_t orc_collect_batch(World* w, int item_id, const float* t4, int64_t n, uint32_t* out_bits,
                           int words_per_query, int nthreads, int mode) {
    const int64_t offs[2] = {0, n};
    return orc_collect_multi(w, 1, &item_id, offs, t4, nullptr, out_bits, words_per_query, nthreads, mode);
}
// DTransformation -> (cos, sin, tx, ty) on the host libm (transformation.rs:131-139)
void orc_compose_batch(const float* rot_tx_ty, int64_t n, float* t4) {
    for (int64_t i = 0; i < n; i++) {
        float r = rot_tx_ty[3 * i];
        t4[4 * i] = cosf(r);
        t4[4 * i + 1] = sinf(r);
        t4[4 * i + 2] = rot_tx_ty[3 * i + 1];
        t4[4 * i + 3] = rot_tx_ty[3 * i + 2];
    }
}

// ---- primitive hooks for unit tests ------------------------------------------------------
int orc_approx_eq(float a, float b) { return approx_eq(a, b) ? 1 : 0; }
int orc_edge_edge(const float* a, const float* b) {
    return edge_edge({{a[0], a[1]}, {a[2], a[3]}}, {{b[0], b[1]}, {b[2], b[3]}}) ? 1 : 0;
}
int orc_rect_edge(const float* r, const float* e) {
    return Rect{r[0], r[1], r[2], r[3]}.collides_with(Edge{{e[0], e[1]}, {e[2], e[3]}}) ? 1 : 0;
}
int orc_circle_rect(const float* c, const float* r) {
    return Circle{{c[0], c[1]}, c[2]}.collides_with(Rect{r[0], r[1], r[2], r[3]}) ? 1 : 0;
}
int orc_circle_edge(const float* c, const float* e) {
    return Circle{{c[0], c[1]}, c[2]}.collides_with(Edge{{e[0], e[1]}, {e[2], e[3]}}) ? 1 : 0;
}
int orc_rect_relation(const float* a, const float* b, int almost) {
    Rect ra{a[0], a[1], a[2], a[3]}, rb{b[0], b[1], b[2], b[3]};
    return (int)(almost ? ra.almost_relation_to(rb) : ra.relation_to(rb));
}
// returns 4-bit mask of colliding quadrants of rect r for the edge
int orc_edge_quadrants(const float* e, const float* r) {
    Rect rr{r[0], r[1], r[2], r[3]};
    Rect qs[4];
    rr.quadrants(qs);
    bool out[4];
    collides_with_quadrants(Edge{{e[0], e[1]}, {e[2], e[3]}}, rr, qs, out);
    return (out[0] ? 1 : 0) | (out[1] ? 2 : 0) | (out[2] ? 4 : 0) | (out[3] ? 8 : 0);
}
int orc_edge_quadrants_individual(const float* e, const float* r) {
    Rect rr{r[0], r[1], r[2], r[3]};
    Rect qs[4];
    rr.quadrants(qs);
    int m = 0;
    for (int i = 0; i < 4; i++)
        if (qs[i].collides_with(Edge{{e[0], e[1]}, {e[2], e[3]}})) m |= 1 << i;
    return m;
}
int orc_pip(const float* xy, int n, float px, float py) {
    SPolygon s;
    s.vertices = to_points(xy, n);
    s.bbox = SPolygon::generate_bounding_box(s.vertices);
    return s.collides_with(Point{px, py}) ? 1 : 0;
}
void orc_transform_points(const float* t4, const float* xy, int n, float* out) {
    Transformation t = from_cstt(t4[0], t4[1], t4[2], t4[3]);
    for (int i = 0; i < n; i++) {
        Point p = transform_point({xy[2 * i], xy[2 * i + 1]}, t);
        out[2 * i] = p.x;
        out[2 * i + 1] = p.y;
    }
}
// import_simple_polygon + SPolygon::new: returns n vertices (CCW), fills xy and info[poi x,y,r, area, diameter]
int orc_import_polygon(const float* xy_in, int n, float* xy_out, float* info) {
    try {
        SPolygon s = import_simple_polygon(to_points(xy_in, n));
        for (int i = 0; i < s.n_vertices(); i++) {
            xy_out[2 * i] = s.vertices[i].x;
            xy_out[2 * i + 1] = s.vertices[i].y;
        }
        info[0] = s.poi.center.x, info[1] = s.poi.center.y, info[2] = s.poi.radius, info[3] = s.area,
        info[4] = s.diameter;
        return s.n_vertices();
    } catch (const std::exception&) {
        return -1;
    }
}
}  // extern "C"
