// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED (see geom.hpp header).
// Restatement of the callers either side of the CDE that the benchmarks need to *build inputs*:
// entities/{item,container,placed_item,layout}.rs, probs/spp/entities/{strip,problem}.rs,
// io/import.rs (item/container import without shape modification: simplify_tolerance = None,
// offset = None, as lbf/benches/util.rs:18-22 passes None for offset and concavity), and an
// LBF-style filler after lbf/src/opt/{search,lbf_spp,loss}.rs + samplers. The filler uses its own
// counter-based PRNG (the reference's SmallRng stream is a third-party detail).
#pragma once
#include "cde.hpp"

namespace orc {

enum class RotKind : int { None = 0, Continuous = 1, Discrete = 2 };  // geo_enums.rs:22-29

struct Item {  // entities/item.rs:12-53
    int id = 0;
    SPolygon shape_orig;         // as imported (CCW), before centering
    DTransformation pre_transform{0, 0, 0};
    SPolygon shape_cd;           // centred + surrogate
    RotKind rot_kind = RotKind::None;
    std::vector<float> rotations;  // radians (Discrete)
    int min_quality = -1;          // -1 = None
    int demand = 1;
};

// io/import.rs:47-98 import_item (ExtShape::SimplePolygon; shape modification disabled)
inline Item import_item(int id, std::vector<Point> pts, const float* orient_deg, int n_orient,
                        int min_quality, int demand, const SPSurrogateConfig& scfg) {
    Item it;
    it.id = id;
    it.shape_orig = import_simple_polygon(std::move(pts));
    Point c = it.shape_orig.centroid();  // io/import.rs:214-217
    it.pre_transform = {0.0f, -c.x, -c.y};
    // geometry/original_shape.rs:23-51 convert_to_internal: transform_clone(pre_transform.compose())
    it.shape_cd = it.shape_orig;
    it.shape_cd.transform(compose(it.pre_transform));
    generate_surrogate(it.shape_cd, scfg);
    if (n_orient < 0) it.rot_kind = RotKind::Continuous;  // io/import.rs:80-89
    else if (n_orient == 0 || (n_orient == 1 && orient_deg[0] == 0.0f)) it.rot_kind = RotKind::None;
    else {
        it.rot_kind = RotKind::Discrete;
        // f32::to_radians: deg * (PI / 180.0) in f32
        const float rads_per_deg = 3.14159265358979323846f / 180.0f;
        for (int i = 0; i < n_orient; i++) it.rotations.push_back(orient_deg[i] * rads_per_deg);
    }
    it.min_quality = min_quality;
    it.demand = demand;
    return it;
}

struct QualityZoneShape {
    int quality;
    SPolygon shape;
};

struct Container {  // entities/container.rs:15-138
    SPolygon outer_cd;
    std::vector<QualityZoneShape> zones;  // grouped by quality ascending, input order inside
    std::shared_ptr<CDEngine> base_cde;
};

// container.rs:28-80 + io/import.rs:100-195 (holes merged into quality 0 after the q0 zones)
inline Container make_container(SPolygon outer, std::vector<QualityZoneShape> zones_in,
                                std::vector<SPolygon> holes, const CDEConfig& cfg) {
    Container c;
    c.outer_cd = std::move(outer);
    for (int q = 0; q < 10; q++) {  // N_QUALITIES = 10, container.rs:88
        for (auto& z : zones_in)
            if (z.quality == q) c.zones.push_back(z);
        if (q == 0)
            for (auto& h : holes) c.zones.push_back({0, h});
    }
    std::vector<Hazard> hazards;
    hazards.push_back({HazardEntity{EntityKind::Exterior}, c.outer_cd, false});
    int cur_q = -1, idx = 0;
    for (auto& z : c.zones) {  // container.rs:121-133 to_hazards: idx enumerates within a quality
        if (z.quality != cur_q) {
            cur_q = z.quality;
            idx = 0;
        }
        HazardEntity e;
        if (z.quality == 0) {
            e.kind = EntityKind::Hole;
            e.id = idx;
        } else {
            e.kind = EntityKind::InferiorQualityZone;
            e.quality = z.quality;
            e.id = idx;
        }
        hazards.push_back({e, z.shape, false});
        idx++;
    }
    c.base_cde = std::make_shared<CDEngine>(c.outer_cd.bbox.inflate_to_square(), std::move(hazards), cfg);
    return c;
}

// probs/spp/entities/strip.rs:40-56 From<Strip> for Container
inline Container strip_container(float width, float height, const CDEConfig& cfg) {
    return make_container(spolygon_from_rect({0.0f, 0.0f, width, height}), {}, {}, cfg);
}

struct PlacedItem {  // entities/placed_item.rs:17-37
    int item_id;
    DTransformation d_transf;
    SPolygon shape;
};

// entities/layout.rs:14-137
struct Layout {
    Container container;
    struct Slot {
        uint32_t version = 0;
        uint32_t next_free = 0;
        std::unique_ptr<PlacedItem> pi;
    };
    std::vector<Slot> placed;  // SlotMap<PItemKey, PlacedItem>
    uint32_t free_head = UINT32_MAX;
    CDEngine cde;

    explicit Layout(Container c) : container(std::move(c)), cde(*container.base_cde) {}

    static uint64_t pk_of(uint32_t idx, uint32_t version) { return ((uint64_t)version << 32) | idx; }
    size_t n_placed() const {
        size_t n = 0;
        for (auto& s : placed) n += s.pi ? 1 : 0;
        return n;
    }
    template <class F>
    void for_each_placed(F f) const {
        for (uint32_t i = 0; i < placed.size(); i++)
            if (placed[i].pi) f(pk_of(i, placed[i].version), *placed[i].pi);
    }
    // layout.rs:73-85
    uint64_t place_item(const Item& item, DTransformation dt) {
        uint32_t idx;
        if (free_head != UINT32_MAX) {
            idx = free_head;
            free_head = placed[idx].next_free;
            placed[idx].version += 1;
        } else {
            idx = (uint32_t)placed.size();
            placed.emplace_back();
            placed[idx].version = 1;
        }
        auto pi = std::make_unique<PlacedItem>();
        pi->item_id = item.id;
        pi->d_transf = dt;
        pi->shape = item.shape_cd;  // placed_item.rs:27-36 transform_clone
        pi->shape.transform(compose(dt));
        uint64_t pk = pk_of(idx, placed[idx].version);
        HazardEntity e;
        e.kind = EntityKind::PlacedItem;
        e.id = item.id;
        e.dt = dt;
        e.pk = pk;
        Hazard hz{e, pi->shape, true};
        placed[idx].pi = std::move(pi);
        cde.register_hazard(std::move(hz));
        return pk;
    }
    // layout.rs:88-100
    PlacedItem remove_item(uint64_t pk) {
        uint32_t idx = (uint32_t)pk, version = (uint32_t)(pk >> 32);
        if (idx >= placed.size() || placed[idx].version != version || !placed[idx].pi)
            throw std::runtime_error("key is not valid anymore");
        PlacedItem pi = std::move(*placed[idx].pi);
        placed[idx].pi.reset();
        placed[idx].version += 1;
        placed[idx].next_free = free_head;
        free_head = idx;
        HazardEntity e;
        e.kind = EntityKind::PlacedItem;
        e.id = pi.item_id;
        e.dt = pi.d_transf;
        e.pk = pk;
        cde.deregister_hazard_by_entity(e);
        return pi;
    }
    // layout.rs:40-49
    void swap_container(Container c) {
        CDESnapshot snap = cde.save();
        container = std::move(c);
        cde = *container.base_cde;
        for (Hazard& h : snap.dynamic_hazards) cde.register_hazard(std::move(h));
    }
    // layout.rs:128-136
    bool is_feasible() const {
        bool ok = true;
        for_each_placed([&](uint64_t pk, const PlacedItem& pi) {
            auto hk = cde.haz_key_from_pi_key(pk);
            if (!hk) throw std::runtime_error("all placed items should be registered in the CDE");
            HazardFilter f;
            f.set(*hk);
            if (cde.detect_poly_collision(pi.shape, f)) ok = false;
        });
        return ok;
    }
};

// Counter-based PRNG (SplitMix64); documented in DESIGN.md. Not the reference's SmallRng.
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    float unit() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }  // [0,1), 24 bits
    float uniform(float lo, float hi) {
        float v = lo + (hi - lo) * unit();
        return v < hi ? v : lo;
    }
    float normal(float mean, float sd) {  // Box-Muller
        float u1 = 1.0f - unit(), u2 = unit();
        return mean + sd * std::sqrt(-2.0f * std::log(u1)) * std::cos(6.28318530717958647692f * u2);
    }
    size_t below(size_t n) { return (size_t)(next() % n); }
};

struct LBFParams {
    int n_samples = 5000;
    float ls_frac = 0.2f;
};

// lbf/src/opt/loss.rs:12-52
struct LBFLoss {
    float x_max, y_max;
    float cost() const { return x_max * 10.0f + y_max; }
};

inline float sample_rotation_uniform(const Item& it, Rng& rng) {  // rotation_distr.rs:32-49
    switch (it.rot_kind) {
        case RotKind::None: return 0.0f;
        case RotKind::Continuous: return rng.uniform(0.0f, 2.0f * 3.14159265358979323846f);
        case RotKind::Discrete: return it.rotations[rng.below(it.rotations.size())];
    }
    return 0.0f;
}

// lbf/src/opt/search.rs:17-106
inline bool lbf_search(const CDEngine& cde, const Item& item, const LBFParams& prm, Rng& rng,
                       const HazardFilter& filter, DTransformation& out) {
    const SPSurrogate& surrogate = item.shape_cd.surrogate;
    SPolygon buffer = item.shape_cd;
    buffer.has_surrogate = false;
    bool has_best = false;
    DTransformation best{0, 0, 0};
    LBFLoss best_cost{0, 0};
    int ls_budget = (int)((float)prm.n_samples * prm.ls_frac);
    int uni_budget = prm.n_samples - ls_budget;
    Rect sbox = cde.bbox();  // uniform_rect_sampler.rs:18-36
    for (int i = 0; i < uni_budget; i++) {
        DTransformation dt;
        dt.rotation = sample_rotation_uniform(item, rng);
        dt.tx = rng.uniform(sbox.x_min, sbox.x_max);
        dt.ty = rng.uniform(sbox.y_min, sbox.y_max);
        Transformation t = compose(dt);
        if (!cde.detect_surrogate_collision(surrogate, t, filter)) {
            buffer.transform_from(item.shape_cd, t);
            LBFLoss cost{buffer.bbox.x_max, buffer.bbox.y_max};
            bool worth = !has_best || cost.cost() < best_cost.cost();
            if (worth && !cde.detect_poly_collision(buffer, filter)) {
                best = dt;
                best_cost = cost;
                has_best = true;
                sbox.x_max = std::fmin(sbox.x_max, cost.cost() / 10.0f);  // loss.rs:38-47
                if (!(sbox.x_min < sbox.x_max)) sbox.x_max = std::nextafter(sbox.x_min, INFINITY);
            }
        }
    }
    if (!has_best) return false;
    // ls_sampler.rs:11-106: normal around the best, exponentially decaying stddev
    Rect cb = cde.bbox();
    float max_dim = std::fmax(cb.width(), cb.height());
    const float sd_t0 = 0.01f * max_dim, sd_t1 = 0.0005f * max_dim;
    const float sd_r0 = 2.0f * 3.14159265358979323846f / 180.0f, sd_r1 = 0.5f * 3.14159265358979323846f / 180.0f;
    float sd_t = sd_t0, sd_r = sd_r0;
    for (int i = 0; i < ls_budget; i++) {
        DTransformation dt;
        dt.rotation = item.rot_kind == RotKind::Continuous ? rng.normal(best.rotation, sd_r)
                      : (item.rot_kind == RotKind::Discrete ? best.rotation : 0.0f);
        dt.tx = rng.normal(best.tx, sd_t);
        dt.ty = rng.normal(best.ty, sd_t);
        Transformation t = compose(dt);
        if (!cde.detect_surrogate_collision(surrogate, t, filter)) {
            buffer.transform_from(item.shape_cd, t);
            LBFLoss cost{buffer.bbox.x_max, buffer.bbox.y_max};
            if (cost.cost() < best_cost.cost() && !cde.detect_poly_collision(buffer, filter)) {
                best = dt;
                best_cost = cost;
            }
        }
        float pct = (float)i / (float)ls_budget;
        sd_t = sd_t0 * std::pow(sd_t1 / sd_t0, pct);
        sd_r = sd_r0 * std::pow(sd_r1 / sd_r0, pct);
    }
    out = best;
    return true;
}

inline HazardFilter item_filter(const CDEngine& cde, const Item& it) {  // lbf_spp.rs:45-66
    if (it.min_quality < 0) return HazardFilter{};
    return filter_from_irrelevant_qzones(it.min_quality, cde.hazards_map);
}

// search.rs:108-115: descending diameter; sorted_by_key is stable
inline std::vector<int> item_placement_order(const std::vector<Item>& items) {
    std::vector<int> order;
    for (auto& it : items) order.push_back(it.id);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return items[a].shape_cd.diameter > items[b].shape_cd.diameter;
    });
    return order;
}

}  // namespace orc
