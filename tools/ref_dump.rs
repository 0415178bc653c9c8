//! Replays the parity cases (tools/export_ref_cases.py -> tests/golden/ref_case_*.json, schema 2) against the REAL jagua-rs
//! crate and reports every query whose verdict or colliding-hazard set differs from what the case file expects (= what
//! this repo's CPU oracle, and hence its GPU kernels, answer). One run pins: quadtree depths 0 / 3 / 5 / 10, surrogate piers
//! (the crate generates its own: fail_fast/piers.rs), continuous rotation, a bin with holes and inferior-quality zones under
//! HazKeyFilter::from_irrelevant_qzones, the self-key filter, and the `unreachable!` of cd_engine.rs:179 (status bit 0x80).
//!
//! NOT compiled in this repo's build image (no cargo / rustc there): written against the reference's public API as it
//! is used by lbf/benches/util.rs, lbf/benches/ci_bench.rs and lbf/src/opt/lbf_bpp.rs. On a machine with a Rust toolchain:
//!
//!     cp tools/ref_dump.rs <jagua-rs checkout>/lbf/examples/ref_dump.rs
//!     cd <jagua-rs checkout>/lbf && for c in <this repo>/tests/golden/ref_case_*.json; do \
//!         cargo run --release --example ref_dump -- $c <this repo> || echo "MISMATCH in $c"; done
//!
//! The cases store f32 values as shortest round-trip decimals; serde_json parses them to f64 and `as f32` is exact.

use std::collections::BTreeSet;
use std::panic::{AssertUnwindSafe, catch_unwind};
use std::path::Path;

use jagua_rs::collision_detection::hazards::HazardEntity;
use jagua_rs::collision_detection::hazards::collector::BasicHazardCollector;
use jagua_rs::collision_detection::hazards::filter::{HazKeyFilter, HazardFilter, NoFilter};
use jagua_rs::collision_detection::{CDEConfig, CDEngine};
use jagua_rs::entities::{Instance, Item, Layout, PItemKey};
use jagua_rs::geometry::DTransformation;
use jagua_rs::geometry::fail_fast::SPSurrogateConfig;
use jagua_rs::geometry::geo_traits::TransformableFrom;
use jagua_rs::io::import::Importer;
use jagua_rs::probs::spp::entities::{SPPlacement, SPProblem};
use jagua_rs::probs::{bpp, spp};
use serde_json::Value;

fn f(v: &Value) -> f32 {
    v.as_f64().expect("number") as f32
}

fn dt(v: &Value) -> DTransformation {
    let a = v.as_array().expect("[rot, tx, ty]");
    DTransformation::new(f(&a[0]), (f(&a[1]), f(&a[2])))
}

struct Tally {
    n: usize,
    bad_detect: usize,
    bad_collect: usize,
}

/// One query set under one filter. Verdict byte: bit0 surrogate, bit1 polygon (evaluated only when bit0 == 0;
/// lbf/benches/ci_bench.rs:107-121), bit7: detect_poly_collision panicked (cd_engine.rs:179).
fn run_set<F: HazardFilter>(cde: &CDEngine, item: &Item, q: &Value, filter: &F, pkeys: &[PItemKey], t: &mut Tally) {
    let surrogate = item.shape_cd.surrogate();
    let mut buffer = item.shape_cd.as_ref().clone();
    let exp_d = q["expect_detect"].as_array().unwrap();
    let exp_c = q["expect_collect"].as_array().unwrap();
    for (k, d) in q["d_transfs"].as_array().unwrap().iter().enumerate() {
        let transf = dt(d).compose();
        let mut verdict = 0u64;
        if cde.detect_surrogate_collision(surrogate, &transf, filter) {
            verdict |= 1;
        } else {
            buffer.transform_from(&item.shape_cd, &transf);
            match catch_unwind(AssertUnwindSafe(|| cde.detect_poly_collision(&buffer, filter))) {
                Ok(true) => verdict |= 2,
                Ok(false) => {}
                Err(_) => verdict |= 0x80,
            }
        }
        if verdict != exp_d[k].as_u64().unwrap() {
            t.bad_detect += 1;
            eprintln!("detect mismatch: item {} query {k}: reference {verdict}, case {}", item.id, exp_d[k]);
        }
        // collect: lbf/benches/ci_bench.rs:56-70 (surrogate and polygon pass share one collector; no filter on this path)
        let mut collector = BasicHazardCollector::with_capacity(cde.hazards_map.len());
        cde.collect_surrogate_collisions(surrogate, &transf, &mut collector);
        buffer.transform_from(&item.shape_cd, &transf);
        cde.collect_poly_collisions(&buffer, &mut collector);
        let got: BTreeSet<i64> = collector
            .iter()
            .map(|(_, e)| match e {
                HazardEntity::Exterior => -1,
                HazardEntity::PlacedItem { pk, .. } => pkeys.iter().position(|x| x == pk).unwrap() as i64,
                HazardEntity::Hole { idx } => 1_000_000 + *idx as i64,
                HazardEntity::InferiorQualityZone { idx, .. } => 2_000_000 + *idx as i64,
            })
            .collect();
        let want: BTreeSet<i64> = exp_c[k].as_array().unwrap().iter().map(|v| v.as_i64().unwrap()).collect();
        if got != want {
            t.bad_collect += 1;
            eprintln!("collect mismatch: item {} query {k}: reference {got:?}, case {want:?}", item.id);
        }
        t.n += 1;
    }
}

fn run_sets(cde: &CDEngine, items: &dyn Fn(usize) -> Item, case: &Value, pkeys: &[PItemKey], t: &mut Tally) {
    for q in case["query_sets"].as_array().unwrap() {
        let item = items(q["item_id"].as_u64().unwrap() as usize);
        let flt = &q["filter"];
        match flt["type"].as_str().unwrap() {
            "none" => run_set(cde, &item, q, &NoFilter, pkeys, t),
            "qzones" => {
                // lbf_bpp.rs:125-127: one filter per item.min_quality, built from the CDE's hazards map
                let filter = HazKeyFilter::from_irrelevant_qzones(flt["required_quality"].as_u64().unwrap() as usize, &cde.hazards_map);
                run_set(cde, &item, q, &filter, pkeys, t)
            }
            "self" => {
                // Layout::is_feasible (entities/layout.rs:128-136): the placed item's own hazard is irrelevant
                let hkey = cde.haz_key_from_pi_key(pkeys[flt["placement"].as_u64().unwrap() as usize]).expect("placement is registered");
                run_set(cde, &item, q, &hkey, pkeys, t)
            }
            other => panic!("unknown filter type {other}"),
        }
    }
}

fn main() {
    std::panic::set_hook(Box::new(|_| {})); // the N3 cases panic on purpose (caught above)
    let args: Vec<String> = std::env::args().collect();
    let case: Value = serde_json::from_reader(std::fs::File::open(&args[1]).expect("case file")).expect("json");
    let repo_root = args.get(2).map(String::as_str).unwrap_or(".");
    assert_eq!(case["schema"].as_u64(), Some(2), "schema 2 case expected");

    // same CDE configuration as the case; shape modification disabled (the case says so)
    let c = &case["cde"];
    let limits = c["n_pole_limits"].as_array().unwrap();
    let mut n_pole_limits = [(0usize, 0.0f32); 3];
    for (k, l) in limits.iter().enumerate() {
        n_pole_limits[k] = (l[0].as_u64().unwrap() as usize, f(&l[1]));
    }
    let cde_config = CDEConfig {
        quadtree_depth: c["quadtree_depth"].as_u64().unwrap() as u8,
        cd_threshold: c["cd_threshold"].as_u64().unwrap() as u8,
        item_surrogate_config: SPSurrogateConfig {
            n_pole_limits,
            n_ff_poles: c["n_ff_poles"].as_u64().unwrap() as usize,
            n_ff_piers: c["n_ff_piers"].as_u64().unwrap() as usize,
        },
    };
    let importer = Importer::new(cde_config, None, None, None);
    let path = Path::new(repo_root).join(case["instance"].as_str().unwrap());
    let mut t = Tally { n: 0, bad_detect: 0, bad_collect: 0 };
    let mut pkeys: Vec<PItemKey> = vec![];

    if case["kind"].as_str() == Some("spp") {
        let ext = lbf::io::read_spp_instance(&path).expect("instance");
        let instance = spp::io::import_instance(&importer, &ext).expect("import");
        // the layout: strip of the exported width, placements in order, then the removals
        let mut problem = SPProblem::new(instance.clone());
        problem.change_strip_width(f(&case["strip_width"]));
        for p in case["placements"].as_array().unwrap() {
            pkeys.push(problem.place_item(SPPlacement { item_id: p["item_id"].as_u64().unwrap() as usize, d_transf: dt(&p["d_transf"]) }));
        }
        for r in case["removed"].as_array().unwrap() {
            problem.remove_item(pkeys[r.as_u64().unwrap() as usize]);
        }
        run_sets(problem.layout.cde(), &|id| instance.item(id).clone(), &case, &pkeys, &mut t);
    } else {
        let ext = lbf::io::read_bpp_instance(&path).expect("instance");
        let instance = bpp::io::import_instance(&importer, &ext).expect("import");
        let bin = &instance.bins[case["bin_index"].as_u64().unwrap() as usize];
        let mut layout = Layout::new(bin.container.clone());
        for p in case["placements"].as_array().unwrap() {
            pkeys.push(layout.place_item(instance.item(p["item_id"].as_u64().unwrap() as usize), dt(&p["d_transf"])));
        }
        for r in case["removed"].as_array().unwrap() {
            layout.remove_item(pkeys[r.as_u64().unwrap() as usize]);
        }
        run_sets(layout.cde(), &|id| instance.item(id).clone(), &case, &pkeys, &mut t);
    }
    println!("{}: {} queries replayed against jagua-rs: {} verdict mismatches, {} hazard-set mismatches",
             case["name"].as_str().unwrap_or("?"), t.n, t.bad_detect, t.bad_collect);
    std::process::exit(if t.bad_detect + t.bad_collect == 0 { 0 } else { 1 });
}
