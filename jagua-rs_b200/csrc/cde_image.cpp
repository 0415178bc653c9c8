// Incremental device image of the host CDE mirror (see cde_host.hpp): a position-stable blob in the format of layout_soa.h
// plus the byte ranges that changed since the previous sync. What the reference does in place on its pointer tree
// (qt_node.rs:38-94 register_hazard / deregister_hazard touch only the nodes the hazard is present in), this does to the
// flattened SoA copy the kernels read: only the touched nodes' records, entry blocks and edge blocks are re-serialised.
#include <algorithm>
#include <cstring>

#include "cde_host.hpp"

namespace jg {

uint64_t HostCDE::next_lineage_ = 1;

static uint32_t align16u(uint32_t x) { return (x + 15u) & ~15u; }

bool HostCDE::Heap::alloc(uint32_t n, Block& b) {
    if (n < free_by_size.size() && !free_by_size[n].empty()) {
        b.off = free_by_size[n].back();
        b.cap = n;
        free_by_size[n].pop_back();
        live += n;
        return true;
    }
    if (top + n > cap) return false;
    b.off = top;
    b.cap = n;
    top += n;
    live += n;
    return true;
}
void HostCDE::Heap::release(Block& b) {
    if (b.cap == 0) return;
    if (b.cap >= free_by_size.size()) free_by_size.resize(b.cap + 1);
    free_by_size[b.cap].push_back(b.off);
    live -= b.cap;
    b = Block{};
}

void HostCDE::image_mark(uint32_t off, uint32_t len) {
    if (len) img_.dirty.push_back(ImageRange{off, len});
}

void HostCDE::image_write_header() {
    JgLayoutHeader& h = img_.hdr;
    h.n_nodes = (uint32_t)nodes_.size();
    h.n_entries = img_.entries.top;
    h.n_edges = img_.edges.top;
    h.n_slots = (uint32_t)slots_.size();
    h.n_haz_verts = img_.verts.top;
    h.exterior_slot = (uint32_t)hkey_exterior();
    h.flags = 0;
    std::memcpy(img_.blob.data(), &h, sizeof(h));
}

// Block sizes are exact: the kernels' working set (and what stage 2 stages into shared memory by TMA) is the sections up to
// their high-water marks, and padding every block to 4 entries / 8 edges made the swim image 38 % larger and the detect
// kernels 7 % slower (0.214 -> 0.228 ms per 1M queries). A node whose hazard count changes simply moves to a block of its new
// size (exact-size free lists); the image is rebuilt - compacted - when turnover has grown a heap to 1.5x its live content.
static uint32_t entry_block(uint32_t n) { return n; }
static uint32_t edge_block(uint32_t n) { return n; }
static uint32_t vert_block(uint32_t n) { return n; }

// node record + its entry and edge blocks; false: a heap is full (the caller rebuilds with larger sections)
bool HostCDE::image_write_node(int node) {
    const QNode& n = nodes_[node];
    const JgLayoutHeader& h = img_.hdr;
    uint8_t* base = img_.blob.data();
    Block& eb = img_.node_entries[node];
    Block& gb = img_.node_edges[node];
    const uint32_t n_ent = n.dead ? 0u : (uint32_t)n.entries.size();
    const uint32_t n_edg = n.dead ? 0u : n.n_active_edges;
    if (n_ent != eb.cap) {
        img_.entries.release(eb);
        if (n_ent && !img_.entries.alloc(entry_block(n_ent), eb)) return false;
    }
    if (n_edg != gb.cap) {
        img_.edges.release(gb);
        if (n_edg && !img_.edges.alloc(edge_block(n_edg), gb)) return false;
    }
    uint32_t n_entire = 0, n_partial = 0;
    if (!n.dead)
        for (const QEntry& q : n.entries) (q.presence == PRES_ENTIRE ? n_entire : n_partial)++;
    int32_t* nm = (int32_t*)(base + h.off_node_meta) + 4 * (size_t)node;
    nm[0] = n.dead ? -1 : n.child_base;
    nm[1] = (int32_t)eb.off;
    nm[2] = (int32_t)(n_entire | (n_partial << 16));
    nm[3] = n.dead ? -2 : n.parent;  // -2: a record of a freed child block (k_grid_build and the validators skip it)
    float* nb = (float*)(base + h.off_node_bbox) + 4 * (size_t)node;
    nb[0] = n.bbox.x_min, nb[1] = n.bbox.y_min, nb[2] = n.bbox.x_max, nb[3] = n.bbox.y_max;
    ((uint32_t*)(base + h.off_node_nactive))[node] = n_edg;
    image_mark(h.off_node_meta + 16u * (uint32_t)node, 16);
    image_mark(h.off_node_bbox + 16u * (uint32_t)node, 16);
    image_mark(h.off_node_nactive + 4u * (uint32_t)node, 4);
    if (n_ent) {
        int32_t* em = (int32_t*)(base + h.off_entry_meta) + 4 * (size_t)eb.off;
        float* ef = (float*)(base + h.off_entry_ffbbox) + 4 * (size_t)eb.off;
        float* ed = (float*)(base + h.off_edges) + 4 * (size_t)gb.off;
        uint32_t gi = 0;
        for (uint32_t i = 0; i < n_ent; i++) {  // already Entire first (entry_insert)
            const QEntry& q = n.entries[i];
            em[4 * i + 0] = (int32_t)q.slot;
            em[4 * i + 1] = (int32_t)(gb.off + gi);
            em[4 * i + 2] = (int32_t)q.edges.size();
            em[4 * i + 3] = q.presence;
            ef[4 * i + 0] = q.ff_bbox.x_min, ef[4 * i + 1] = q.ff_bbox.y_min, ef[4 * i + 2] = q.ff_bbox.x_max, ef[4 * i + 3] = q.ff_bbox.y_max;
            for (const Seg& s : q.edges) {
                ed[4 * gi + 0] = s.sx, ed[4 * gi + 1] = s.sy, ed[4 * gi + 2] = s.ex, ed[4 * gi + 3] = s.ey;
                gi++;
            }
        }
        image_mark(h.off_entry_meta + 16u * eb.off, 16u * n_ent);
        image_mark(h.off_entry_ffbbox + 16u * eb.off, 16u * n_ent);
        image_mark(h.off_edges + 16u * gb.off, 16u * gi);
    }
    return true;
}

// this node's share of the grid premise (after ALL dirty records were written: the check reads the parent's record too)
void HostCDE::image_check_node(int node) {
    if (!img_.axes_ok) return;
    const JgLayoutHeader& h = img_.hdr;
    const uint8_t* base = img_.blob.data();
    const uint8_t bad = grid_node_ok((const int32_t*)(base + h.off_node_meta), (const float*)(base + h.off_node_bbox), (uint32_t)node,
                                     (uint32_t)nodes_.size(), cfg_.quadtree_depth, img_.xs, img_.ys) ? 0 : 1;
    img_.n_bad += (int)bad - (int)img_.node_bad[node];
    img_.node_bad[node] = bad;
}

bool HostCDE::image_write_slot(uint32_t s) {
    const JgLayoutHeader& h = img_.hdr;
    uint8_t* base = img_.blob.data();
    if (s >= img_.cap_slots) return false;
    if (s >= img_.slot_verts.size()) img_.slot_verts.resize(s + 1);
    Block& vb = img_.slot_verts[s];
    const Slot& sl = slots_[s];
    const uint32_t nv = sl.live ? (uint32_t)sl.hz.n_verts() : 0u;
    if (nv != vb.cap) {
        img_.verts.release(vb);
        if (nv && !img_.verts.alloc(vert_block(nv), vb)) return false;
    }
    int32_t* hm = (int32_t*)(base + h.off_haz_meta) + 4 * (size_t)s;
    float* hb = (float*)(base + h.off_haz_bbox) + 4 * (size_t)s;
    float* hp = (float*)(base + h.off_haz_poi) + 4 * (size_t)s;
    std::memset(hm, 0, 16), std::memset(hb, 0, 16), std::memset(hp, 0, 16);
    if (sl.live) {
        const HostHazard& hz = sl.hz;
        hm[0] = hz.scope(), hm[1] = (int32_t)vb.off, hm[2] = hz.n_verts(), hm[3] = hz.kind;
        hb[0] = hz.bbox.x_min, hb[1] = hz.bbox.y_min, hb[2] = hz.bbox.x_max, hb[3] = hz.bbox.y_max;
        hp[0] = hz.poi_x, hp[1] = hz.poi_y;
        std::memcpy((float*)(base + h.off_haz_verts) + 2 * (size_t)vb.off, hz.xy.data(), hz.xy.size() * sizeof(float));
        image_mark(h.off_haz_verts + 8u * vb.off, 8u * nv);
    }
    ((uint32_t*)(base + h.off_haz_version))[s] = sl.version;
    image_mark(h.off_haz_meta + 16u * s, 16);
    image_mark(h.off_haz_bbox + 16u * s, 16);
    image_mark(h.off_haz_poi + 16u * s, 16);
    image_mark(h.off_haz_version + 4u * s, 4);
    return true;
}

// A fresh image with room to grow: every section at a multiple of what the tree needs now.
void HostCDE::image_rebuild() {
    uint32_t n_ent = 0, n_edg = 0, n_vert = 0;
    for (const QNode& n : nodes_)
        if (!n.dead) n_ent += entry_block((uint32_t)n.entries.size()), n_edg += edge_block(n.n_active_edges);
    for (const Slot& s : slots_)
        if (s.live) n_vert += vert_block((uint32_t)s.hz.n_verts());
    Image& I = img_;
    I = Image{};
    // a quadtree of depth <= 5 has at most 1365 nodes: the node sections never move; deeper trees get twice their size
    const uint32_t full5 = 1365;
    I.cap_nodes = cfg_.quadtree_depth <= 5 ? full5 : std::max<uint32_t>(2 * (uint32_t)nodes_.size(), 4096);
    if (I.cap_nodes < nodes_.size()) I.cap_nodes = 2 * (uint32_t)nodes_.size();
    I.cap_slots = std::max<uint32_t>(64, 2 * (uint32_t)slots_.size());
    I.entries.cap = std::max<uint32_t>(1024, n_ent + n_ent / 2 + 256);
    I.edges.cap = std::max<uint32_t>(4096, n_edg + n_edg / 2 + 1024);
    I.verts.cap = std::max<uint32_t>(1024, n_vert + n_vert / 2 + 512);
    JgLayoutHeader& h = I.hdr;
    std::memset(&h, 0, sizeof(h));
    h.magic = JG_LAYOUT_MAGIC;
    h.version = JG_LAYOUT_VERSION;
    h.root_level = (uint32_t)cfg_.quadtree_depth;
    h.cd_threshold = (uint32_t)cfg_.cd_threshold;
    const Rect rb = nodes_[0].bbox;
    h.cde_bbox[0] = rb.x_min, h.cde_bbox[1] = rb.y_min, h.cde_bbox[2] = rb.x_max, h.cde_bbox[3] = rb.y_max;
    uint32_t off = align16u((uint32_t)sizeof(JgLayoutHeader));
    auto section = [&](uint32_t& field, uint32_t bytes) {
        field = off;
        off = align16u(off + bytes);
    };
    section(h.off_node_meta, I.cap_nodes * 16);
    section(h.off_node_bbox, I.cap_nodes * 16);
    section(h.off_node_nactive, I.cap_nodes * 4);
    section(h.off_entry_meta, I.entries.cap * 16);
    section(h.off_entry_ffbbox, I.entries.cap * 16);
    section(h.off_edges, I.edges.cap * 16);
    section(h.off_haz_meta, I.cap_slots * 16);
    section(h.off_haz_bbox, I.cap_slots * 16);
    section(h.off_haz_poi, I.cap_slots * 16);
    section(h.off_haz_verts, I.verts.cap * 8);
    section(h.off_haz_version, I.cap_slots * 4);
    h.total_bytes = off;
    I.blob.assign(off, 0);
    I.node_entries.assign(I.cap_nodes, Block{});
    I.node_edges.assign(I.cap_nodes, Block{});
    I.slot_verts.assign(I.cap_slots, Block{});
    I.lineage = next_lineage_++;
    I.seq = 0;
    I.valid = true;
    I.axes_ok = nodes_.size() <= 0xFFFEu && grid_axes_ok(h.cde_bbox, cfg_.quadtree_depth, I.xs, I.ys);
    I.node_bad.assign(I.cap_nodes, 0);
    I.n_bad = 0;
    for (int n = 0; n < (int)nodes_.size(); n++) image_write_node(n);  // (cannot fail: the heaps were sized for this)
    for (int n = 0; n < (int)nodes_.size(); n++) image_check_node(n);
    for (uint32_t s = 0; s < slots_.size(); s++) image_write_slot(s);
    I.dirty.clear();  // a new lineage is taken whole
}

const std::vector<uint8_t>& HostCDE::image_sync(std::vector<ImageRange>& ranges, uint64_t& lineage, uint64_t& seq) {
    ranges.clear();
    bool rebuilt = false;
    if (!img_.valid || nodes_.size() > img_.cap_nodes || slots_.size() > img_.cap_slots) {
        image_rebuild();
        rebuilt = true;
    } else {
        bool ok = true;
        for (int n : dirty_nodes_) ok = ok && image_write_node(n);
        if (ok)
            for (int n : dirty_nodes_) image_check_node(n);
        for (uint32_t s : dirty_slots_) ok = ok && image_write_slot(s);
        // a full heap, or heaps grown to twice their live content by block turnover: start over (compacts)
        if (!ok || img_.entries.top > img_.entries.live + img_.entries.live / 2 + 256 || img_.edges.top > img_.edges.live + img_.edges.live / 2 + 1024) {
            image_rebuild();
            rebuilt = true;
        }
    }
    for (int n : dirty_nodes_)
        if ((size_t)n < node_dirty_.size()) node_dirty_[n] = 0;
    dirty_nodes_.clear();
    dirty_slots_.clear();
    image_write_header();
    // the grid premise of the detect / collect grid kernels, evaluated here so that an upload needs no read-back
    const bool grid_ok = img_.axes_ok && img_.n_bad == 0 && nodes_.size() <= 0xFFFEu;
    const uint32_t flags = JG_LAYOUT_FLAG_GRID_CHECKED | (grid_ok ? JG_LAYOUT_FLAG_GRID_OK : 0u);
    img_.hdr.flags = flags;
    std::memcpy(img_.blob.data() + offsetof(JgLayoutHeader, flags), &flags, sizeof(flags));
    img_.seq++;
    lineage = img_.lineage;
    seq = img_.seq;
    if (!rebuilt) {
        image_mark(0, (uint32_t)sizeof(JgLayoutHeader));
        // merge: sort by offset, join ranges that touch or lie within 256 bytes of each other (fewer, larger copies)
        std::vector<ImageRange>& d = img_.dirty;
        std::sort(d.begin(), d.end(), [](const ImageRange& a, const ImageRange& b) { return a.off < b.off; });
        for (const ImageRange& r : d) {
            if (!ranges.empty() && r.off <= ranges.back().off + ranges.back().len + 256u) {
                const uint32_t end = std::max(ranges.back().off + ranges.back().len, r.off + r.len);
                ranges.back().len = end - ranges.back().off;
            } else {
                ranges.push_back(r);
            }
        }
    }
    img_.dirty.clear();
    return img_.blob;
}

}  // namespace jg
