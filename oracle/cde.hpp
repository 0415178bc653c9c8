// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see geom.hpp header).
// Restatement of collision_detection/{cd_engine.rs, quadtree/*.rs, hazards/*.rs}: same recursion,
// same iteration orders, same comparisons. Hazard identity is a slot index with a version, like
// slotmap's (idx, version) HazKey.
#pragma once
#include <array>
#include <memory>
#include <optional>

#include "spoly.hpp"

namespace orc {

// hazards/hazard.rs:36-62
enum class EntityKind : int { PlacedItem = 0, Exterior = 1, Hole = 2, InferiorQualityZone = 3 };
struct HazardEntity {
    EntityKind kind = EntityKind::Exterior;
    int64_t id = 0;        // PlacedItem.id / Hole.idx / InferiorQualityZone.idx
    int64_t quality = 0;   // InferiorQualityZone.quality
    uint64_t pk = 0;       // PlacedItem.pk (PItemKey as u64)
    DTransformation dt{0, 0, 0};
    bool operator==(const HazardEntity& o) const {
        if (kind != o.kind) return false;
        switch (kind) {
            case EntityKind::PlacedItem: return id == o.id && pk == o.pk && dt == o.dt;
            case EntityKind::Exterior: return true;
            case EntityKind::Hole: return id == o.id;
            case EntityKind::InferiorQualityZone: return id == o.id && quality == o.quality;
        }
        return false;
    }
    GeoPosition scope() const {  // hazard.rs:55-62
        return kind == EntityKind::Exterior ? GeoPosition::Exterior : GeoPosition::Interior;
    }
};

struct Hazard {  // hazard.rs:16-33
    HazardEntity entity;
    SPolygon shape;
    bool dynamic;
};

// Minimal slotmap: u32 index + u32 version, LIFO free list (slotmap 1.0 semantics that matter here:
// keys of removed entries never compare equal to later keys of the same slot; iteration is in
// slot order).
struct HazKey {
    uint32_t idx = 0, version = 0;
    bool operator==(const HazKey& o) const { return idx == o.idx && version == o.version; }
    uint64_t as_u64() const { return ((uint64_t)version << 32) | idx; }
    static HazKey from_u64(uint64_t v) { return {(uint32_t)v, (uint32_t)(v >> 32)}; }
};
struct HazardMap {
    struct Slot {
        uint32_t version = 0;  // odd = occupied
        uint32_t next_free = 0;
        std::unique_ptr<Hazard> value;
    };
    std::vector<Slot> slots;
    uint32_t free_head = UINT32_MAX;
    size_t n = 0;
    HazKey insert(Hazard h) {
        uint32_t idx;
        if (free_head != UINT32_MAX) {
            idx = free_head;
            free_head = slots[idx].next_free;
            slots[idx].version += 1;
        } else {
            idx = (uint32_t)slots.size();
            slots.emplace_back();
            slots[idx].version = 1;
        }
        slots[idx].value = std::make_unique<Hazard>(std::move(h));
        n++;
        return {idx, slots[idx].version};
    }
    bool contains(HazKey k) const {
        return k.idx < slots.size() && slots[k.idx].version == k.version && slots[k.idx].value;
    }
    const Hazard& at(HazKey k) const {
        if (!contains(k)) throw std::runtime_error("invalid HazKey");
        return *slots[k.idx].value;
    }
    std::optional<Hazard> remove(HazKey k) {
        if (!contains(k)) return std::nullopt;
        Hazard h = std::move(*slots[k.idx].value);
        slots[k.idx].value.reset();
        slots[k.idx].version += 1;
        slots[k.idx].next_free = free_head;
        free_head = k.idx;
        n--;
        return h;
    }
    template <class F>
    void for_each(F f) const {
        for (uint32_t i = 0; i < slots.size(); i++)
            if (slots[i].value) f(HazKey{i, slots[i].version}, *slots[i].value);
    }
    size_t capacity() const { return slots.size(); }
    HazardMap() = default;
    HazardMap(const HazardMap& o) : free_head(o.free_head), n(o.n) {
        slots.resize(o.slots.size());
        for (size_t i = 0; i < o.slots.size(); i++) {
            slots[i].version = o.slots[i].version;
            slots[i].next_free = o.slots[i].next_free;
            if (o.slots[i].value) slots[i].value = std::make_unique<Hazard>(*o.slots[i].value);
        }
    }
    HazardMap& operator=(const HazardMap& o) {
        if (this != &o) {
            HazardMap tmp(o);
            std::swap(slots, tmp.slots);
            free_head = tmp.free_head;
            n = tmp.n;
        }
        return *this;
    }
};

// hazards/filter.rs:7-76 — every filter of the reference is a set of HazKeys; the oracle keeps it
// as a bitset over slot indices (keys are only ever taken from the live hazards_map).
struct HazardFilter {
    std::vector<uint32_t> bits;
    bool is_irrelevant(HazKey k) const {
        uint32_t w = k.idx >> 5;
        return w < bits.size() && ((bits[w] >> (k.idx & 31)) & 1u);
    }
    void set(HazKey k) {
        uint32_t w = k.idx >> 5;
        if (w >= bits.size()) bits.resize(w + 1, 0);
        bits[w] |= 1u << (k.idx & 31);
    }
};
// filter.rs:21-41 from_irrelevant_qzones: irrelevant iff InferiorQualityZone.quality < required
inline HazardFilter filter_from_irrelevant_qzones(int64_t required_quality, const HazardMap& map) {
    HazardFilter f;
    map.for_each([&](HazKey k, const Hazard& h) {
        if (h.entity.kind == EntityKind::InferiorQualityZone && h.entity.quality < required_quality) f.set(k);
    });
    return f;
}

// hazards/collector.rs:7-67 — BasicHazardCollector = SecondaryMap<HazKey, HazardEntity>; the
// oracle's collector is the order-normalised bitset over slot indices.
struct HazardCollector {
    std::vector<uint32_t> bits;
    explicit HazardCollector(size_t cap) : bits((cap + 31) / 32, 0) {}
    bool contains_key(HazKey k) const { return (bits[k.idx >> 5] >> (k.idx & 31)) & 1u; }
    void insert(HazKey k) { bits[k.idx >> 5] |= 1u << (k.idx & 31); }
    void clear() { std::fill(bits.begin(), bits.end(), 0u); }
};

// quadtree/qt_partial_hazard.rs:7-72
struct QTHazPartial {
    std::vector<Edge> edges;
    Rect ff_bbox;
    static QTHazPartial from_entire_shape(const SPolygon& s) {
        QTHazPartial p;
        for (int i = 0; i < s.n_vertices(); i++) p.edges.push_back(s.edge(i));
        p.ff_bbox = s.bbox;
        return p;
    }
    static QTHazPartial from_parent(const QTHazPartial& parent, std::vector<Edge> restricted) {
        QTHazPartial p;
        if (parent.edges.size() == restricted.size()) {
            p.ff_bbox = parent.ff_bbox;
        } else {
            float inf = std::numeric_limits<float>::infinity();
            float x_min = inf, y_min = inf, x_max = -inf, y_max = -inf;
            for (const Edge& e : restricted) {
                x_min = std::fmin(std::fmin(x_min, e.start.x), e.end.x);
                y_min = std::fmin(std::fmin(y_min, e.start.y), e.end.y);
                x_max = std::fmax(std::fmax(x_max, e.start.x), e.end.x);
                y_max = std::fmax(std::fmax(y_max, e.start.y), e.end.y);
            }
            if (x_min < x_max && y_min < y_max) p.ff_bbox = {x_min, y_min, x_max, y_max};
            else p.ff_bbox = parent.ff_bbox;
        }
        p.edges = std::move(restricted);
        return p;
    }
    template <class T>
    bool collides_with(const T& entity) const {  // qt_partial_hazard.rs:67-72
        if (!collides(entity, ff_bbox)) return false;
        for (const Edge& e : edges)
            if (collides(entity, e)) return true;
        return false;
    }
};

enum class Presence : int { None = 0, Partial = 1, Entire = 2 };  // qt_hazard.rs:26-33

// quadtree/qt_hazard.rs:13-175
struct QTHazard {
    Rect qt_bbox;
    HazKey hkey;
    HazardEntity entity;
    Presence presence;
    QTHazPartial partial;  // valid iff presence == Partial

    static QTHazard from_root(const Rect& root_bbox, const Hazard& haz, HazKey hkey) {
        return {root_bbox, hkey, haz.entity, Presence::Partial, QTHazPartial::from_entire_shape(haz.shape)};
    }
    size_t n_edges() const { return presence == Presence::Partial ? partial.edges.size() : 0; }

    static constexpr int NEIGHBORS[4][2] = {{1, 3}, {0, 2}, {1, 3}, {0, 2}};  // rect.rs:136

    std::array<QTHazard, 4> constrict(const Rect quadrants[4], const HazardMap& haz_map) const {
        std::array<QTHazard, 4> out;
        if (presence == Presence::Entire) {
            // qt_hazard.rs:58: four clones of self (qt_bbox stays the parent's; it is never read)
            for (int i = 0; i < 4; i++) out[i] = *this;
            return out;
        }
        const SPolygon& haz_shape = haz_map.at(hkey).shape;
        int enclosed = -1;
        for (int i = 0; i < 4; i++)
            if (haz_shape.bbox.relation_to(quadrants[i]) == GeoRelation::Enclosed) {
                enclosed = i;
                break;
            }
        if (enclosed >= 0) {  // qt_hazard.rs:70-85
            for (int i = 0; i < 4; i++) {
                out[i].qt_bbox = quadrants[i];
                out[i].hkey = hkey;
                out[i].entity = entity;
                if (i == enclosed) {
                    out[i].presence = Presence::Partial;
                    out[i].partial = partial;
                } else {
                    out[i].presence = Presence::None;
                }
            }
            return out;
        }
        bool resolved[4] = {false, false, false, false};
        for (int i = 0; i < 4; i++) {  // qt_hazard.rs:91-109
            std::vector<Edge> colliding;
            for (const Edge& e : partial.edges)
                if (quadrants[i].collides_with(e)) colliding.push_back(e);
            if (!colliding.empty()) {
                out[i] = {quadrants[i], hkey, entity, Presence::Partial,
                          QTHazPartial::from_parent(partial, std::move(colliding))};
                resolved[i] = true;
            }
        }
        for (int i = 0; i < 4; i++) {  // qt_hazard.rs:116-160
            if (resolved[i]) continue;
            bool none_neighbor = false, entire_neighbor = false;
            for (int k = 0; k < 2; k++) {
                int nb = NEIGHBORS[i][k];
                if (!resolved[nb]) continue;
                if (out[nb].presence == Presence::None) none_neighbor = true;
                if (out[nb].presence == Presence::Entire) entire_neighbor = true;
            }
            Presence p;
            if (none_neighbor && entire_neighbor)
                throw std::runtime_error("unreachable: None and Entire neighbours (qt_hazard.rs:136)");
            else if (none_neighbor) p = Presence::None;
            else if (entire_neighbor) p = Presence::Entire;
            else {
                bool colliding = haz_shape.collides_with(quadrants[i].centroid());
                if (entity.scope() == GeoPosition::Interior) p = colliding ? Presence::Entire : Presence::None;
                else p = !colliding ? Presence::Entire : Presence::None;
            }
            out[i] = {quadrants[i], hkey, entity, p, {}};
            resolved[i] = true;
        }
        return out;
    }
};

// quadtree/qt_hazard_vec.rs:13-120
struct QTHazardVec {
    std::vector<QTHazard> hazards;
    size_t n_active_edges = 0;
    static int sortkey(const QTHazard& h) { return (int)h.presence; }
    void add(QTHazard haz) {
        // binary_search_by(|probe| order_by_descending_strength(probe, &haz)) — a partition point in
        // a vector sorted by descending presence. Which index is chosen among equal-presence
        // entries depends on std's binary search; verdicts and hazard sets do not.
        // Restated after core::slice::binary_search_by (Rust 1.8x): size-halving loop, then one
        // final compare.
        size_t size = hazards.size();
        size_t pos = 0;
        if (size > 0) {
            size_t base = 0;
            auto cmp = [&](size_t i) {  // Ordering of probe vs haz under descending strength
                int a = sortkey(hazards[i]), b = sortkey(haz);
                return a > b ? -1 : (a < b ? 1 : 0);  // reverse of a.cmp(b)
            };
            while (size > 1) {
                size_t half = size / 2, mid = base + half;
                if (cmp(mid) != 1) base = mid;  // base = if cmp == Greater { base } else { mid }
                size -= half;
            }
            int c = cmp(base);
            pos = (c == 0) ? base : base + (c == -1 ? 1 : 0);
        }
        n_active_edges += haz.n_edges();
        hazards.insert(hazards.begin() + pos, std::move(haz));
    }
    std::optional<QTHazard> remove(HazKey k) {
        for (size_t i = 0; i < hazards.size(); i++)
            if (hazards[i].hkey == k) {
                QTHazard h = std::move(hazards[i]);
                hazards.erase(hazards.begin() + i);
                n_active_edges -= h.n_edges();
                return h;
            }
        return std::nullopt;
    }
    const QTHazard* strongest(const HazardFilter& f) const {  // qt_hazard_vec.rs:60-63
        for (const QTHazard& h : hazards)
            if (!f.is_irrelevant(h.hkey)) return &h;
        return nullptr;
    }
    bool no_partial_hazards() const {
        for (const QTHazard& h : hazards)
            if (h.presence == Presence::Partial) return false;
        return true;
    }
};

// quadtree/qt_traits.rs:144-177: 0 = no hit, else half of e2: 1 First (u > 0.5), 2 Second, 3 Halfway
inline int edge_intersection_half(const Edge& e1, const Edge& e2) {
    ORC_FLOPS(31);
    float x1 = e1.start.x, y1 = e1.start.y, x2 = e1.end.x, y2 = e1.end.y;
    float x3 = e2.start.x, y3 = e2.start.y, x4 = e2.end.x, y4 = e2.end.y;
    float t_nom = (x2 - x4) * (y4 - y3) - (y2 - y4) * (x4 - x3);
    float t_denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    float u_nom = (x2 - x4) * (y2 - y1) - (y2 - y4) * (x2 - x1);
    float u_denom = (x2 - x1) * (y4 - y3) - (y2 - y1) * (x4 - x3);
    if (t_denom == 0.0f || u_denom == 0.0f) return 0;
    float t = t_nom / t_denom, u = u_nom / u_denom;
    if (t >= 0.0f && t <= 1.0f && u >= 0.0f && u <= 1.0f) return u > 0.5f ? 1 : (u < 0.5f ? 2 : 3);
    return 0;
}

// quadtree/qt_traits.rs:10-13 default (Circle): four entity∩Rect tests
inline void collides_with_quadrants(const Circle& c, const Rect&, const Rect qs[4], bool out[4]) {
    for (int i = 0; i < 4; i++) out[i] = c.collides_with(qs[i]);
}
// quadtree/qt_traits.rs:19-106 (Edge). c[q]: -1 = None, 0 = Some(false), 1 = Some(true)
inline void collides_with_quadrants(const Edge& e, const Rect& r, const Rect qs[4], bool out[4]) {
    float ex_min = e.x_min(), ex_max = e.x_max(), ey_min = e.y_min(), ey_max = e.y_max();
    ORC_FLOPS(4);
    int c[4];
    for (int i = 0; i < 4; i++) {
        const Rect& q = qs[i];
        bool x_no = std::fmax(ex_min, q.x_min) > std::fmin(ex_max, q.x_max);
        bool y_no = std::fmax(ey_min, q.y_min) > std::fmin(ey_max, q.y_max);
        ORC_FLOPS(6);
        if (x_no || y_no) c[i] = 0;
        else {
            ORC_FLOPS(8);
            if (q.collides_with(e.start) || q.collides_with(e.end)) c[i] = 1;
            else c[i] = -1;
        }
    }
    if (c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[3] >= 0) {
        for (int i = 0; i < 4; i++) out[i] = c[i] == 1;
        return;
    }
    Point ctr = r.centroid();
    Point k[4];
    r.corners(k);  // rect.rs:164-184 sides: top k0->k1, left k1->k2, bottom k2->k3, right k3->k0
    Edge top{k[0], k[1]}, left{k[1], k[2]}, bottom{k[2], k[3]}, right{k[3], k[0]};
    Edge h_bisect{{r.x_min, ctr.y}, {r.x_max, ctr.y}};
    Edge v_bisect{{ctr.x, r.y_min}, {ctr.x, r.y_max}};
    // qt_traits.rs:110-140 half_intersect
    auto half_intersect = [&](const Edge& e2, std::initializer_list<int> fst, std::initializer_list<int> sec) {
        bool any_none = false;
        for (int q : fst) any_none = any_none || c[q] < 0;
        for (int q : sec) any_none = any_none || c[q] < 0;
        if (!any_none) return;
        int loc = edge_intersection_half(e, e2);
        if (loc == 0) return;
        if (loc == 1 || loc == 3)
            for (int q : fst) c[q] = 1;
        if (loc == 2 || loc == 3)
            for (int q : sec) c[q] = 1;
    };
    half_intersect(left, {1}, {2});
    half_intersect(right, {3}, {0});
    half_intersect(top, {0}, {1});
    half_intersect(bottom, {2}, {3});
    half_intersect(h_bisect, {1, 2}, {0, 3});
    half_intersect(v_bisect, {2, 3}, {0, 1});
    for (int i = 0; i < 4; i++) out[i] = c[i] == 1;
}

// quadtree/qt_node.rs:14-192
struct QTNode {
    uint8_t level;
    Rect bbox;
    std::unique_ptr<std::array<QTNode, 4>> children;
    QTHazardVec hazards;
    uint8_t cd_threshold;

    QTNode() : level(0), bbox{0, 0, 0, 0}, cd_threshold(0) {}
    QTNode(uint8_t lvl, Rect bb, uint8_t thr) : level(lvl), bbox(bb), cd_threshold(thr) {}
    QTNode(const QTNode& o)
        : level(o.level), bbox(o.bbox), hazards(o.hazards), cd_threshold(o.cd_threshold) {
        if (o.children) children = std::make_unique<std::array<QTNode, 4>>(*o.children);
    }
    QTNode& operator=(const QTNode& o) {
        if (this != &o) {
            level = o.level;
            bbox = o.bbox;
            hazards = o.hazards;
            cd_threshold = o.cd_threshold;
            children = o.children ? std::make_unique<std::array<QTNode, 4>>(*o.children) : nullptr;
        }
        return *this;
    }
    QTNode(QTNode&&) = default;
    QTNode& operator=(QTNode&&) = default;

    void constrict_and_register_to_children(const QTHazard& h, const HazardMap& map) {
        Rect cb[4];
        for (int i = 0; i < 4; i++) cb[i] = (*children)[i].bbox;
        std::array<QTHazard, 4> ch = h.constrict(cb, map);
        for (int i = 0; i < 4; i++)
            if (ch[i].presence != Presence::None) (*children)[i].register_hazard(std::move(ch[i]), map);
    }
    // qt_node.rs:38-81
    void register_hazard(QTHazard new_haz, const HazardMap& map) {
        if (!children && level > 0 && new_haz.presence == Presence::Partial) {
            Rect q[4];
            bbox.quadrants(q);
            children = std::make_unique<std::array<QTNode, 4>>();
            for (int i = 0; i < 4; i++) (*children)[i] = QTNode(level - 1, q[i], cd_threshold);
            for (const QTHazard& h : hazards.hazards) constrict_and_register_to_children(h, map);
        }
        if (children) constrict_and_register_to_children(new_haz, map);
        hazards.add(std::move(new_haz));
    }
    // qt_node.rs:83-94
    void deregister_hazard(HazKey k) {
        bool modified = hazards.remove(k).has_value();
        if (modified) {
            if (hazards.no_partial_hazards()) children.reset();
            else if (children)
                for (QTNode& c : *children) c.deregister_hazard(k);
        }
    }
    // qt_node.rs:99-147. Returns the colliding QTHazard or nullptr.
    template <class T>
    const QTHazard* collides(const T& entity, const HazardFilter& filter) const {
        const QTHazard* strongest = hazards.strongest(filter);
        if (!strongest) return nullptr;
        if (strongest->presence == Presence::None) return nullptr;
        if (strongest->presence == Presence::Entire) return strongest;
        if (children) {
            Rect qs[4];
            for (int i = 0; i < 4; i++) qs[i] = (*children)[i].bbox;
            bool cq[4];
            collides_with_quadrants(entity, bbox, qs, cq);
            for (int i = 0; i < 4; i++)
                if (cq[i]) {
                    const QTHazard* r = (*children)[i].collides(entity, filter);
                    if (r) return r;
                }
            return nullptr;
        }
        for (const QTHazard& hz : hazards.hazards) {
            if (filter.is_irrelevant(hz.hkey)) continue;
            if (hz.presence == Presence::Entire)
                throw std::runtime_error("unreachable: should have been handled above (qt_node.rs:137)");
            if (hz.presence == Presence::Partial && hz.partial.collides_with(entity)) return &hz;
        }
        return nullptr;
    }
    // qt_node.rs:151-191
    template <class T>
    void collect_collisions(const T& entity, HazardCollector& collector) const {
        bool perform_cd_now = hazards.n_active_edges <= (size_t)cd_threshold;
        if (children && !perform_cd_now) {
            Rect qs[4];
            for (int i = 0; i < 4; i++) qs[i] = (*children)[i].bbox;
            bool cq[4];
            collides_with_quadrants(entity, bbox, qs, cq);
            for (int i = 0; i < 4; i++)
                if (cq[i]) (*children)[i].collect_collisions(entity, collector);
        } else {
            for (const QTHazard& hz : hazards.hazards) {
                if (collector.contains_key(hz.hkey)) continue;
                if (hz.presence == Presence::Entire) collector.insert(hz.hkey);
                else if (hz.presence == Presence::Partial && hz.partial.collides_with(entity))
                    collector.insert(hz.hkey);
            }
        }
    }
};

struct CDEConfig {  // cd_engine.rs:346-353
    uint8_t quadtree_depth = 5;
    uint8_t cd_threshold = 16;
    SPSurrogateConfig item_surrogate_config;
};

struct CDESnapshot {  // cd_engine.rs:357-359
    std::vector<Hazard> dynamic_hazards;
};

struct PanicInReference : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// collision_detection/cd_engine.rs:22-342
struct CDEngine {
    QTNode quadtree;
    HazardMap hazards_map;
    CDEConfig config;
    HazKey hkey_exterior;

    CDEngine(Rect bbox, std::vector<Hazard> static_hazards, CDEConfig cfg)
        : quadtree(cfg.quadtree_depth, bbox, cfg.cd_threshold), config(cfg) {
        for (Hazard& h : static_hazards) {
            HazKey k = hazards_map.insert(std::move(h));
            quadtree.register_hazard(QTHazard::from_root(quadtree.bbox, hazards_map.at(k), k), hazards_map);
        }
        bool found = false;
        hazards_map.for_each([&](HazKey k, const Hazard& h) {
            if (!found && h.entity.kind == EntityKind::Exterior) {
                hkey_exterior = k;
                found = true;
            }
        });
        if (!found) throw std::runtime_error("No exterior hazard registered in the CDE");
    }
    Rect bbox() const { return quadtree.bbox; }

    HazKey register_hazard(Hazard hazard) {  // cd_engine.rs:59-69
        HazKey k = hazards_map.insert(std::move(hazard));
        quadtree.register_hazard(QTHazard::from_root(bbox(), hazards_map.at(k), k), hazards_map);
        return k;
    }
    Hazard deregister_hazard_by_entity(const HazardEntity& ent) {  // cd_engine.rs:72-85
        bool found = false;
        HazKey key;
        hazards_map.for_each([&](HazKey k, const Hazard& h) {
            if (!found && h.entity == ent) {
                key = k;
                found = true;
            }
        });
        if (!found) throw std::runtime_error("Cannot deregister hazard that is not registered");
        quadtree.deregister_hazard(key);
        return *hazards_map.remove(key);
    }
    Hazard deregister_hazard_by_key(HazKey k) {  // cd_engine.rs:87-96
        auto h = hazards_map.remove(k);
        if (!h) throw std::runtime_error("Cannot deregister hazard that is not registered");
        quadtree.deregister_hazard(k);
        return std::move(*h);
    }
    CDESnapshot save() const {  // cd_engine.rs:98-106
        CDESnapshot s;
        hazards_map.for_each([&](HazKey, const Hazard& h) {
            if (h.dynamic) s.dynamic_hazards.push_back(h);
        });
        return s;
    }
    void restore(const CDESnapshot& snap) {  // cd_engine.rs:109-149
        std::vector<std::pair<HazKey, HazardEntity>> to_remove;
        hazards_map.for_each([&](HazKey k, const Hazard& h) {
            if (h.dynamic) to_remove.push_back({k, h.entity});
        });
        std::vector<Hazard> to_add;
        for (const Hazard& h : snap.dynamic_hazards) {
            size_t pos = to_remove.size();
            for (size_t i = 0; i < to_remove.size(); i++)
                if (to_remove[i].second == h.entity) {
                    pos = i;
                    break;
                }
            if (pos < to_remove.size()) {  // swap_remove
                to_remove[pos] = to_remove.back();
                to_remove.pop_back();
            } else {
                to_add.push_back(h);
            }
        }
        for (auto& kv : to_remove) deregister_hazard_by_key(kv.first);
        for (Hazard& h : to_add) register_hazard(std::move(h));
    }

    // cd_engine.rs:314-327
    const QTNode* get_virtual_root(const Rect& bb) const {
        const QTNode* v = &quadtree;
        while (v->children) {
            const QTNode* next = nullptr;
            for (const QTNode& c : *v->children)
                if (c.bbox.relation_to(bb) == GeoRelation::Surrounding) {
                    next = &c;
                    break;
                }
            if (!next) break;
            v = next;
        }
        return v;
    }

    // cd_engine.rs:231-255
    bool detect_containment_collision(const SPolygon& shape, const SPolygon& haz_shape,
                                      const HazardEntity& ent) const {
        GeoRelation rel = haz_shape.bbox.almost_relation_to(shape.bbox);
        bool contained;
        switch (rel) {
            case GeoRelation::Surrounding: contained = haz_shape.collides_with(shape.poi.center); break;
            case GeoRelation::Enclosed: contained = shape.collides_with(haz_shape.poi.center); break;
            default: contained = false;
        }
        return ent.scope() == GeoPosition::Interior ? contained : !contained;
    }

    // cd_engine.rs:160-197. Throws PanicInReference where the reference hits `unreachable!` (N3).
    bool detect_poly_collision(const SPolygon& shape, const HazardFilter& filter) const {
        if (bbox().relation_to(shape.bbox) != GeoRelation::Surrounding) return true;
        const QTNode* v = get_virtual_root(shape.bbox);
        for (int i = 0; i < shape.n_vertices(); i++)
            if (v->collides(shape.edge(i), filter)) return true;
        for (const QTHazard& qh : v->hazards.hazards) {
            if (qh.presence == Presence::Entire)
                throw PanicInReference("unreachable: Entire hazard in the virtual root (cd_engine.rs:179)");
            if (qh.presence == Presence::Partial && !filter.is_irrelevant(qh.hkey)) {
                if (detect_containment_collision(shape, hazards_map.at(qh.hkey).shape, qh.entity)) return true;
            }
        }
        return false;
    }
    // cd_engine.rs:204-223
    bool detect_surrogate_collision(const SPSurrogate& s, const Transformation& t,
                                    const HazardFilter& filter) const {
        for (int i = 0; i < s.n_ff_poles(); i++)
            if (quadtree.collides(s.poles[i].transformed(t), filter)) return true;
        for (const Edge& pier : s.piers)
            if (quadtree.collides(pier.transformed(t), filter)) return true;
        return false;
    }
    // cd_engine.rs:261-289
    void collect_poly_collisions(const SPolygon& shape, HazardCollector& col) const {
        if (bbox().relation_to(shape.bbox) != GeoRelation::Surrounding) col.insert(hkey_exterior);
        const QTNode* v = get_virtual_root(shape.bbox);
        for (int i = 0; i < shape.n_vertices(); i++) v->collect_collisions(shape.edge(i), col);
        for (const QTHazard& qh : v->hazards.hazards) {
            if (qh.presence != Presence::Partial) continue;
            if (!col.contains_key(qh.hkey)) {
                if (detect_containment_collision(shape, hazards_map.at(qh.hkey).shape, qh.entity))
                    col.insert(qh.hkey);
            }
        }
    }
    // cd_engine.rs:296-310
    void collect_surrogate_collisions(const SPSurrogate& s, const Transformation& t,
                                      HazardCollector& col) const {
        for (int i = 0; i < s.n_ff_poles(); i++) quadtree.collect_collisions(s.poles[i].transformed(t), col);
        for (const Edge& pier : s.piers) quadtree.collect_collisions(pier.transformed(t), col);
    }
    // cd_engine.rs:333-341
    std::optional<HazKey> haz_key_from_pi_key(uint64_t pk) const {
        std::optional<HazKey> r;
        hazards_map.for_each([&](HazKey k, const Hazard& h) {
            if (!r && h.entity.kind == EntityKind::PlacedItem && h.entity.pk == pk) r = k;
        });
        return r;
    }
};

}  // namespace orc
