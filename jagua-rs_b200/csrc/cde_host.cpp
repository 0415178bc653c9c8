// Host mirror of CDEngine's update path (see cde_host.hpp). Compile with -ffp-contract=off.
#include "cde_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>

namespace jg {

static const int QUADRANT_NEIGHBORS[4][2] = {{1, 3}, {0, 2}, {1, 3}, {0, 2}};  // rect.rs:136

HostCDE::HostCDE(Rect bbox, JgConfig cfg) : cfg_(cfg) {
    QNode root;
    root.level = (uint8_t)cfg.quadtree_depth;
    root.bbox = bbox;
    nodes_.push_back(std::move(root));
}

// ---- arena ----------------------------------------------------------------------------------
void HostCDE::touch_node(int node) {
    if ((size_t)node >= node_dirty_.size()) node_dirty_.resize(nodes_.size(), 0);
    if (!node_dirty_[node]) {
        node_dirty_[node] = 1;
        dirty_nodes_.push_back(node);
    }
}
void HostCDE::touch_slot(uint32_t slot) { dirty_slots_.push_back(slot); }

int HostCDE::alloc_children(int parent, const Rect& bbox, uint8_t child_level) {
    int base;
    if (!free_blocks_.empty()) {
        base = free_blocks_.back();
        free_blocks_.pop_back();
    } else {
        base = (int)nodes_.size();
        for (int i = 0; i < 4; i++) nodes_.emplace_back();
    }
    for (int i = 0; i < 4; i++) {  // qt_node.rs:65-69: one child per quadrant, in quadrant order
        QNode& c = nodes_[base + i];
        c.level = child_level;
        c.bbox = rect_quadrant(bbox, i);
        c.child_base = -1;
        c.parent = parent;
        c.dead = false;
        c.entries.clear();
        c.n_active_edges = 0;
        touch_node(base + i);
    }
    return base;
}

void HostCDE::free_children(int child_base) {
    for (int i = 0; i < 4; i++) {
        QNode& c = nodes_[child_base + i];
        if (c.child_base >= 0) free_children(c.child_base);
        c.child_base = -1;
        c.dead = true;
        c.entries.clear();
        c.entries.shrink_to_fit();
        c.n_active_edges = 0;
        touch_node(child_base + i);
    }
    free_blocks_.push_back(child_base);
}

// qt_hazard_vec.rs:24-46: the vector stays sorted Entire -> Partial. The position inside a presence
// class is an artefact of std's binary search in the reference and never observable through
// verdicts or hazard sets; here a new entry goes to the end of its class.
void HostCDE::entry_insert(QNode& n, QEntry&& e) {
    n.n_active_edges += (uint32_t)e.edges.size();
    if (e.presence == PRES_ENTIRE) {
        size_t pos = 0;
        while (pos < n.entries.size() && n.entries[pos].presence == PRES_ENTIRE) pos++;
        n.entries.insert(n.entries.begin() + pos, std::move(e));
    } else {
        n.entries.push_back(std::move(e));
    }
}

// ---- constrict (qt_hazard.rs:48-167) ----------------------------------------------------------
void HostCDE::constrict(const QEntry& e, const Rect quads[4], QEntry out[4]) const {
    for (int i = 0; i < 4; i++) {
        out[i] = QEntry{};
        out[i].slot = e.slot;
        out[i].version = e.version;
    }
    if (e.presence == PRES_ENTIRE) {  // present in the entire parent -> entire in all four
        for (int i = 0; i < 4; i++) out[i].presence = PRES_ENTIRE;
        return;
    }
    const HostHazard& hz = slots_[e.slot].hz;

    // a quadrant that encloses the whole hazard takes the parent's partial as is (qt_hazard.rs:65-85)
    for (int i = 0; i < 4; i++) {
        if (rect_relation(hz.bbox, quads[i]) == REL_ENCLOSED) {
            out[i].presence = PRES_PARTIAL;
            out[i].ff_bbox = e.ff_bbox;
            out[i].edges = e.edges;
            return;  // the other three stay PRES_NONE
        }
    }

    bool resolved[4] = {false, false, false, false};
    for (int i = 0; i < 4; i++) {  // qt_hazard.rs:91-109: edges of the parent that touch quadrant i
        std::vector<Seg> kept;
        for (const Seg& s : e.edges)
            if (rect_seg(quads[i], s)) kept.push_back(s);
        if (kept.empty()) continue;
        // qt_partial_hazard.rs:20-60 from_parent: tight ff_bbox unless all edges kept / degenerate
        Rect ff = e.ff_bbox;
        if (kept.size() != e.edges.size()) {
            const float inf = std::numeric_limits<float>::infinity();
            float x0 = inf, y0 = inf, x1 = -inf, y1 = -inf;
            for (const Seg& s : kept) {
                x0 = fminf(fminf(x0, s.sx), s.ex);
                y0 = fminf(fminf(y0, s.sy), s.ey);
                x1 = fmaxf(fmaxf(x1, s.sx), s.ex);
                y1 = fmaxf(fmaxf(y1, s.sy), s.ey);
            }
            if (x0 < x1 && y0 < y1) ff = Rect{x0, y0, x1, y1};
        }
        out[i].presence = PRES_PARTIAL;
        out[i].ff_bbox = ff;
        out[i].edges = std::move(kept);
        resolved[i] = true;
    }
    // qt_hazard.rs:116-160: quadrants without edges are Entire or None; copy a resolved neighbour's
    // Entire/None (in index order, later quadrants see earlier results), else test the centroid.
    for (int i = 0; i < 4; i++) {
        if (resolved[i]) continue;
        bool nb_none = false, nb_entire = false;
        for (int k = 0; k < 2; k++) {
            int nb = QUADRANT_NEIGHBORS[i][k];
            if (!resolved[nb]) continue;
            nb_none = nb_none || out[nb].presence == PRES_NONE;
            nb_entire = nb_entire || out[nb].presence == PRES_ENTIRE;
        }
        uint8_t p;
        if (nb_none && nb_entire) {
            // `unreachable!` in the reference (qt_hazard.rs:136); surfaced by the caller
            const_cast<HostCDE*>(this)->last_error = "reference panic: quadrant with both None and Entire neighbours";
            p = PRES_NONE;
        } else if (nb_none) p = PRES_NONE;
        else if (nb_entire) p = PRES_ENTIRE;
        else {
            const float cx = (quads[i].x_min + quads[i].x_max) / 2.0f;
            const float cy = (quads[i].y_min + quads[i].y_max) / 2.0f;
            const bool inside = poly_point<false>(hz.xy.data(), hz.n_verts(), hz.bbox, cx, cy, Xf{});
            p = (hz.scope() == 1 ? inside : !inside) ? PRES_ENTIRE : PRES_NONE;
        }
        out[i].presence = p;
        resolved[i] = true;
    }
}

void HostCDE::constrict_into_children(int node, const QEntry& e) {
    const int cb = nodes_[node].child_base;
    Rect quads[4];
    for (int i = 0; i < 4; i++) quads[i] = nodes_[cb + i].bbox;
    QEntry ch[4];
    constrict(e, quads, ch);
    for (int i = 0; i < 4; i++)
        if (ch[i].presence != PRES_NONE) node_register(cb + i, std::move(ch[i]));
}

// qt_node.rs:38-81
void HostCDE::node_register(int node, QEntry&& e) {
    if (nodes_[node].child_base < 0 && nodes_[node].level > 0 && e.presence == PRES_PARTIAL) {
        const int cb = alloc_children(node, nodes_[node].bbox, (uint8_t)(nodes_[node].level - 1));
        nodes_[node].child_base = cb;
        // copy: registering into children never touches this node's entries, but keep it simple
        const size_t n_prev = nodes_[node].entries.size();
        for (size_t k = 0; k < n_prev; k++) constrict_into_children(node, nodes_[node].entries[k]);
    }
    if (nodes_[node].child_base >= 0) constrict_into_children(node, e);
    entry_insert(nodes_[node], std::move(e));
    touch_node(node);
}

// qt_node.rs:83-94
void HostCDE::node_deregister(int node, uint32_t slot, uint32_t version) {
    QNode& n = nodes_[node];
    size_t pos = n.entries.size();
    for (size_t k = 0; k < n.entries.size(); k++)
        if (n.entries[k].slot == slot && n.entries[k].version == version) {
            pos = k;
            break;
        }
    if (pos == n.entries.size()) return;  // not modified
    touch_node(node);
    n.n_active_edges -= (uint32_t)n.entries[pos].edges.size();
    n.entries.erase(n.entries.begin() + pos);
    bool any_partial = false;
    for (const QEntry& q : n.entries) any_partial = any_partial || q.presence == PRES_PARTIAL;
    if (!any_partial) {
        if (n.child_base >= 0) free_children(n.child_base);
        n.child_base = -1;
    } else if (n.child_base >= 0) {
        const int cb = n.child_base;
        for (int i = 0; i < 4; i++) node_deregister(cb + i, slot, version);
    }
}

// ---- CDEngine surface ---------------------------------------------------------------------------
uint64_t HostCDE::register_hazard(HostHazard&& hz) {  // cd_engine.rs:59-69
    // simple_polygon.rs:122-133 generate_bounding_box
    float x0 = std::numeric_limits<float>::max(), y0 = x0;
    float x1 = std::numeric_limits<float>::lowest(), y1 = x1;
    for (int i = 0; i < hz.n_verts(); i++) {
        x0 = fminf(x0, hz.xy[2 * i]);
        y0 = fminf(y0, hz.xy[2 * i + 1]);
        x1 = fmaxf(x1, hz.xy[2 * i]);
        y1 = fmaxf(y1, hz.xy[2 * i + 1]);
    }
    hz.bbox = Rect{x0, y0, x1, y1};

    uint32_t idx;
    if (free_head_ != UINT32_MAX) {  // slotmap: LIFO reuse, occupied versions are odd
        idx = free_head_;
        free_head_ = slots_[idx].next_free;
        slots_[idx].version += 1;
    } else {
        idx = (uint32_t)slots_.size();
        slots_.emplace_back();
        slots_[idx].version = 1;
    }
    slots_[idx].live = true;
    slots_[idx].hz = std::move(hz);
    n_live_++;

    QEntry e;  // qt_hazard.rs:36-44 from_root + qt_partial_hazard.rs:15-19 from_entire_shape
    e.slot = idx;
    e.version = slots_[idx].version;
    e.presence = PRES_PARTIAL;
    e.ff_bbox = slots_[idx].hz.bbox;
    const HostHazard& h = slots_[idx].hz;
    e.edges.reserve(h.n_verts());
    for (int i = 0; i < h.n_verts(); i++) e.edges.push_back(h.edge(i));
    node_register(0, std::move(e));
    touch_slot(idx);
    epoch++;
    return make_key(idx, slots_[idx].version);
}

bool HostCDE::deregister_hazard(uint64_t hkey) {  // cd_engine.rs:87-96
    const uint32_t idx = (uint32_t)hkey, version = (uint32_t)(hkey >> 32);
    if (idx >= slots_.size() || !slots_[idx].live || slots_[idx].version != version) {
        last_error = "Cannot deregister hazard that is not registered";
        return false;
    }
    node_deregister(0, idx, version);
    slots_[idx].live = false;
    slots_[idx].hz = HostHazard{};
    slots_[idx].version += 1;
    slots_[idx].next_free = free_head_;
    free_head_ = idx;
    n_live_--;
    touch_slot(idx);
    epoch++;
    return true;
}

uint64_t HostCDE::hkey_from_pk(uint64_t pk) const {  // cd_engine.rs:333-341
    for (uint32_t i = 0; i < slots_.size(); i++)
        if (slots_[i].live && slots_[i].hz.kind == JG_HAZ_PLACED_ITEM && slots_[i].hz.pk == pk)
            return make_key(i, slots_[i].version);
    return 0;
}

uint64_t HostCDE::hkey_from_entity(int kind, int64_t id, int64_t quality, uint64_t pk) const {  // cd_engine.rs:73-78: first match in slot order
    HostHazard probe;
    probe.kind = kind, probe.id = id, probe.quality = quality, probe.pk = pk;
    for (uint32_t i = 0; i < slots_.size(); i++)
        if (slots_[i].live && slots_[i].hz.same_entity(probe)) return make_key(i, slots_[i].version);
    return 0;
}

void HostCDE::swap_container(const HostCDE& base) {  // entities/layout.rs:40-50
    std::vector<HostHazard> dyn;  // self.cde.save(): dynamic hazards in slot order (cd_engine.rs:98-106)
    for (const Slot& s : slots_)
        if (s.live && s.hz.dynamic) dyn.push_back(s.hz);
    const uint64_t ep = epoch;
    *this = base;  // self.cde = container.base_cde.clone()
    invalidate_image();
    snapshots_.clear();
    last_error.clear();
    epoch = ep + 1;
    for (HostHazard& h : dyn) register_hazard(std::move(h));
}

uint64_t HostCDE::hkey_exterior() const {  // cd_engine.rs:44-48: first Exterior in slot order
    for (uint32_t i = 0; i < slots_.size(); i++)
        if (slots_[i].live && slots_[i].hz.kind == JG_HAZ_EXTERIOR) return make_key(i, slots_[i].version);
    return 0;
}

int HostCDE::save() {  // cd_engine.rs:98-106
    std::vector<HostHazard> dyn;
    for (const Slot& s : slots_)
        if (s.live && s.hz.dynamic) dyn.push_back(s.hz);
    snapshots_.push_back(std::move(dyn));
    return (int)snapshots_.size() - 1;
}

bool HostCDE::restore(int snapshot_id) {  // cd_engine.rs:109-149: diff on dynamic hazards
    if (snapshot_id < 0 || snapshot_id >= (int)snapshots_.size()) {
        last_error = "unknown snapshot";
        return false;
    }
    const std::vector<HostHazard>& snap = snapshots_[snapshot_id];
    std::vector<uint32_t> to_remove;  // slot indices of the current dynamic hazards
    for (uint32_t i = 0; i < slots_.size(); i++)
        if (slots_[i].live && slots_[i].hz.dynamic) to_remove.push_back(i);
    std::vector<const HostHazard*> to_add;
    for (const HostHazard& h : snap) {
        size_t pos = to_remove.size();
        for (size_t k = 0; k < to_remove.size(); k++)
            if (slots_[to_remove[k]].hz.same_entity(h)) {
                pos = k;
                break;
            }
        if (pos < to_remove.size()) {  // Vec::swap_remove
            to_remove[pos] = to_remove.back();
            to_remove.pop_back();
        } else {
            to_add.push_back(&h);
        }
    }
    for (uint32_t idx : to_remove) deregister_hazard(make_key(idx, slots_[idx].version));
    for (const HostHazard* h : to_add) register_hazard(HostHazard(*h));
    return true;
}

void HostCDE::filter_qzones(int required_quality, uint32_t* bits, int words) const {  // filter.rs:21-41
    for (int i = 0; i < words; i++) bits[i] = 0;
    for (uint32_t i = 0; i < slots_.size(); i++)
        if (slots_[i].live && slots_[i].hz.kind == JG_HAZ_QZONE && slots_[i].hz.quality < required_quality &&
            (int)(i >> 5) < words)
            bits[i >> 5] |= 1u << (i & 31);
}

// ---- flatten -------------------------------------------------------------------------------------
static uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// The grid premise of csrc/detect_grid.cuh in two parts: what depends on the root box alone (the nested midpoints and the
// one-correction cell lookup), and what every node must satisfy (its box is the grid's, a leaf record fits). The incremental
// image re-checks only the nodes it re-serialises.
bool grid_axes_ok(const float root_bbox[4], int D, float xs[33], float ys[33]) {
    if (D < 0 || D > 5) return false;
    const int G = 1 << D;
    const float* nb = root_bbox;
    xs[0] = nb[0], xs[G] = nb[2], ys[0] = nb[1], ys[G] = nb[3];
    if (!(nb[2] > nb[0] && nb[3] > nb[1])) return false;
    for (int step = G; step >= 2; step >>= 1)
        for (int a = 0; a < G; a += step) {  // rect.rs:248-253 centroid, as k_grid_build nests it
            xs[a + step / 2] = (xs[a] + xs[a + step]) / 2.0f;
            ys[a + step / 2] = (ys[a] + ys[a + step]) / 2.0f;
        }
    const float inv_cw = (float)G / (nb[2] - nb[0]), inv_ch = (float)G / (nb[3] - nb[1]);
    for (int k = 0; k <= G; k++) {  // premise of the one-correction cell lookup
        const float fx = std::floor((xs[k] - xs[0]) * inv_cw), fy = std::floor((ys[k] - ys[0]) * inv_ch);
        if (!(fx >= (float)(k - 1) && fx <= (float)k && fy >= (float)(k - 1) && fy <= (float)k)) return false;
        if (k > 0 && !(xs[k] > xs[k - 1] && ys[k] > ys[k - 1])) return false;
    }
    return true;
}
bool grid_node_ok(const int32_t* nm, const float* nb, uint32_t n, uint32_t n_nodes, int D, const float* xs, const float* ys) {
    if (nm[4 * n + 3] == -2) return true;  // a record of a freed child block (incremental image): not part of the tree
    int lv = 0, ci = 0, cj = 0;
    uint32_t cur = n;
    while (cur > 0 && lv <= D) {
        const int q = (int)((cur - 1) & 3u);
        ci |= ((q == 0 || q == 3) ? 1 : 0) << lv;
        cj |= ((q == 0 || q == 1) ? 1 : 0) << lv;
        lv++;
        const int32_t par = nm[4 * cur + 3];
        if (par < 0 || (uint32_t)par >= n_nodes || nm[4 * par] != (int32_t)(((cur - 1) & ~3u) + 1)) return false;
        cur = (uint32_t)par;
    }
    if (lv > D || cur != 0) return false;
    const int s = D - lv;
    if (!(nb[4 * n] == xs[ci << s] && nb[4 * n + 2] == xs[(ci + 1) << s] && nb[4 * n + 1] == ys[cj << s] &&
          nb[4 * n + 3] == ys[(cj + 1) << s]))
        return false;
    const int n_entire = nm[4 * n + 2] & 0xFFFF, n_partial = (nm[4 * n + 2] >> 16) & 0xFFFF;
    const bool has_children = nm[4 * n] >= 0;
    if (n_entire == 0 && n_partial > 0) {
        if (!has_children && (lv != D || n_partial > 255 || nm[4 * n + 1] + n_entire >= (1 << 24))) return false;  // leaf record
        if (has_children && lv >= D) return false;
    }
    return true;
}

bool blob_grid_ok(const uint8_t* blob) {
    JgLayoutHeader h;
    std::memcpy(&h, blob, sizeof(h));
    const int D = (int)h.root_level;
    if (D < 0 || D > 5 || h.n_nodes == 0 || h.n_nodes > 0xFFFEu) return false;
    const int32_t* nm = (const int32_t*)(blob + h.off_node_meta);
    const float* nb = (const float*)(blob + h.off_node_bbox);
    float xs[33], ys[33];
    if (!grid_axes_ok(nb, D, xs, ys)) return false;
    for (uint32_t n = 0; n < h.n_nodes; n++)
        if (!grid_node_ok(nm, nb, n, h.n_nodes, D, xs, ys)) return false;
    return true;
}

std::vector<uint8_t> HostCDE::flatten() const {
    // breadth-first renumbering: flat index 0 = root, the 4 children of a node are contiguous
    std::vector<int> order{0};   // arena index by flat index
    std::vector<int> parent{-1};
    std::vector<int> child_flat;
    for (size_t f = 0; f < order.size(); f++) {
        const QNode& n = nodes_[order[f]];
        if (n.child_base >= 0) {
            child_flat.push_back((int)order.size());
            for (int i = 0; i < 4; i++) {
                order.push_back(n.child_base + i);
                parent.push_back((int)f);
            }
        } else {
            child_flat.push_back(-1);
        }
    }
    const uint32_t n_nodes = (uint32_t)order.size();
    uint32_t n_entries = 0, n_edges = 0, n_haz_verts = 0;
    for (int a : order) {
        n_entries += (uint32_t)nodes_[a].entries.size();
        n_edges += nodes_[a].n_active_edges;
    }
    const uint32_t n_slots = (uint32_t)slots_.size();
    for (const Slot& s : slots_)
        if (s.live) n_haz_verts += (uint32_t)s.hz.n_verts();

    JgLayoutHeader h;
    std::memset(&h, 0, sizeof(h));
    h.magic = JG_LAYOUT_MAGIC;
    h.version = JG_LAYOUT_VERSION;
    h.n_nodes = n_nodes;
    h.n_entries = n_entries;
    h.n_edges = n_edges;
    h.n_slots = n_slots;
    h.n_haz_verts = n_haz_verts;
    h.root_level = (uint32_t)cfg_.quadtree_depth;
    h.cd_threshold = (uint32_t)cfg_.cd_threshold;
    h.exterior_slot = (uint32_t)hkey_exterior();
    const Rect rb = nodes_[0].bbox;
    h.cde_bbox[0] = rb.x_min, h.cde_bbox[1] = rb.y_min, h.cde_bbox[2] = rb.x_max, h.cde_bbox[3] = rb.y_max;
    uint32_t off = align16((uint32_t)sizeof(JgLayoutHeader));
    auto section = [&](uint32_t& field, uint32_t bytes) {
        field = off;
        off = align16(off + bytes);
    };
    section(h.off_node_meta, n_nodes * 16);
    section(h.off_node_bbox, n_nodes * 16);
    section(h.off_node_nactive, n_nodes * 4);
    section(h.off_entry_meta, n_entries * 16);
    section(h.off_entry_ffbbox, n_entries * 16);
    section(h.off_edges, n_edges * 16);
    section(h.off_haz_meta, n_slots * 16);
    section(h.off_haz_bbox, n_slots * 16);
    section(h.off_haz_poi, n_slots * 16);
    section(h.off_haz_verts, n_haz_verts * 8);
    section(h.off_haz_version, n_slots * 4);
    h.total_bytes = off;

    std::vector<uint8_t> blob(off, 0);
    std::memcpy(blob.data(), &h, sizeof(h));
    int32_t* node_meta = (int32_t*)(blob.data() + h.off_node_meta);
    float* node_bbox = (float*)(blob.data() + h.off_node_bbox);
    uint32_t* node_nactive = (uint32_t*)(blob.data() + h.off_node_nactive);
    int32_t* entry_meta = (int32_t*)(blob.data() + h.off_entry_meta);
    float* entry_ff = (float*)(blob.data() + h.off_entry_ffbbox);
    float* edges = (float*)(blob.data() + h.off_edges);
    int32_t* haz_meta = (int32_t*)(blob.data() + h.off_haz_meta);
    float* haz_bbox = (float*)(blob.data() + h.off_haz_bbox);
    float* haz_poi = (float*)(blob.data() + h.off_haz_poi);
    float* haz_verts = (float*)(blob.data() + h.off_haz_verts);

    uint32_t ei = 0, gi = 0;
    for (uint32_t f = 0; f < n_nodes; f++) {
        const QNode& n = nodes_[order[f]];
        uint32_t n_entire = 0, n_partial = 0;
        for (const QEntry& q : n.entries) (q.presence == PRES_ENTIRE ? n_entire : n_partial)++;
        node_meta[4 * f + 0] = child_flat[f];
        node_meta[4 * f + 1] = (int32_t)ei;
        node_meta[4 * f + 2] = (int32_t)(n_entire | (n_partial << 16));
        node_meta[4 * f + 3] = parent[f];
        node_bbox[4 * f + 0] = n.bbox.x_min, node_bbox[4 * f + 1] = n.bbox.y_min;
        node_bbox[4 * f + 2] = n.bbox.x_max, node_bbox[4 * f + 3] = n.bbox.y_max;
        node_nactive[f] = n.n_active_edges;
        for (const QEntry& q : n.entries) {  // already Entire first
            entry_meta[4 * ei + 0] = (int32_t)q.slot;
            entry_meta[4 * ei + 1] = (int32_t)gi;
            entry_meta[4 * ei + 2] = (int32_t)q.edges.size();
            entry_meta[4 * ei + 3] = q.presence;
            entry_ff[4 * ei + 0] = q.ff_bbox.x_min, entry_ff[4 * ei + 1] = q.ff_bbox.y_min;
            entry_ff[4 * ei + 2] = q.ff_bbox.x_max, entry_ff[4 * ei + 3] = q.ff_bbox.y_max;
            for (const Seg& s : q.edges) {
                edges[4 * gi + 0] = s.sx, edges[4 * gi + 1] = s.sy, edges[4 * gi + 2] = s.ex, edges[4 * gi + 3] = s.ey;
                gi++;
            }
            ei++;
        }
    }
    uint32_t* haz_version = (uint32_t*)(blob.data() + h.off_haz_version);
    uint32_t vi = 0;
    for (uint32_t s = 0; s < n_slots; s++) {
        haz_version[s] = slots_[s].version;  // odd = occupied (slotmap); a dead slot keeps the even version it was freed with
        if (!slots_[s].live) continue;  // dead slot: vert_count 0
        const HostHazard& hz = slots_[s].hz;
        haz_meta[4 * s + 0] = hz.scope();
        haz_meta[4 * s + 1] = (int32_t)vi;
        haz_meta[4 * s + 2] = hz.n_verts();
        haz_meta[4 * s + 3] = hz.kind;
        haz_bbox[4 * s + 0] = hz.bbox.x_min, haz_bbox[4 * s + 1] = hz.bbox.y_min;
        haz_bbox[4 * s + 2] = hz.bbox.x_max, haz_bbox[4 * s + 3] = hz.bbox.y_max;
        haz_poi[4 * s + 0] = hz.poi_x, haz_poi[4 * s + 1] = hz.poi_y;
        std::memcpy(haz_verts + 2 * vi, hz.xy.data(), hz.xy.size() * sizeof(float));
        vi += (uint32_t)hz.n_verts();
    }
    // the grid premise of the detect / collect grid kernels, evaluated here so that an upload needs no read-back
    const uint32_t flags = JG_LAYOUT_FLAG_GRID_CHECKED | (blob_grid_ok(blob.data()) ? JG_LAYOUT_FLAG_GRID_OK : 0u);
    std::memcpy(blob.data() + offsetof(JgLayoutHeader, flags), &flags, sizeof(flags));
    return blob;
}

// ---- canonical dump (tests compare it word for word with the oracle's) -------------------------------
void HostCDE::dump_node(int node, std::vector<uint32_t>& out) const {
    const QNode& n = nodes_[node];
    auto pf = [&](float f) { out.push_back(f32_bits(f)); };
    out.push_back(n.level);
    pf(n.bbox.x_min), pf(n.bbox.y_min), pf(n.bbox.x_max), pf(n.bbox.y_max);
    out.push_back(n.child_base >= 0 ? 1 : 0);
    out.push_back((uint32_t)n.entries.size());
    out.push_back(n.n_active_edges);
    std::vector<const QEntry*> es;
    for (const QEntry& q : n.entries) es.push_back(&q);
    std::stable_sort(es.begin(), es.end(), [](const QEntry* a, const QEntry* b) {
        if (a->presence != b->presence) return a->presence > b->presence;
        return a->slot < b->slot;
    });
    for (const QEntry* q : es) {
        out.push_back(q->slot);
        out.push_back(q->presence);
        if (q->presence == PRES_PARTIAL) {
            pf(q->ff_bbox.x_min), pf(q->ff_bbox.y_min), pf(q->ff_bbox.x_max), pf(q->ff_bbox.y_max);
            out.push_back((uint32_t)q->edges.size());
            for (const Seg& s : q->edges) pf(s.sx), pf(s.sy), pf(s.ex), pf(s.ey);
        } else {
            for (int i = 0; i < 5; i++) out.push_back(0);
        }
    }
    if (n.child_base >= 0)
        for (int i = 0; i < 4; i++) dump_node(n.child_base + i, out);
}

std::vector<uint32_t> HostCDE::tree_dump() const {
    std::vector<uint32_t> out;
    dump_node(0, out);
    return out;
}

}  // namespace jg
