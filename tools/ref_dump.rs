//! Replays a parity case (tools/export_ref_case.py -> tests/golden/ref_case_*.json) against the REAL jagua-rs crate and
//! reports every query whose verdict or colliding-hazard set differs from what the case file expects (= what this
//! repo's CPU oracle, and hence its GPU kernels, answer).
//!
//! NOT compiled in this repo's build image (no cargo / rustc there): written against the reference's public API as it
//! is used by lbf/benches/util.rs and lbf/benches/ci_bench.rs. On a machine with a Rust toolchain:
//!
//!     cp tools/ref_dump.rs <jagua-rs checkout>/lbf/examples/ref_dump.rs
//!     cd <jagua-rs checkout>/lbf && cargo run --release --example ref_dump -- <this repo>/tests/golden/ref_case_swim.json <this repo>
//!
//! The case stores f32 values as shortest round-trip decimals; serde_json parses them to f64 and `as f32` is exact.

use std::collections::BTreeSet;
use std::path::Path;

use jagua_rs::collision_detection::CDEConfig;
use jagua_rs::collision_detection::hazards::HazardEntity;
use jagua_rs::collision_detection::hazards::collector::BasicHazardCollector;
use jagua_rs::collision_detection::hazards::filter::NoFilter;
use jagua_rs::entities::Instance;
use jagua_rs::geometry::DTransformation;
use jagua_rs::geometry::fail_fast::SPSurrogateConfig;
use jagua_rs::geometry::geo_traits::TransformableFrom;
use jagua_rs::io::import::Importer;
use jagua_rs::probs::spp;
use jagua_rs::probs::spp::entities::{SPPlacement, SPProblem};
use serde_json::Value;

fn f(v: &Value) -> f32 {
    v.as_f64().expect("number") as f32
}

fn dt(v: &Value) -> DTransformation {
    let a = v.as_array().expect("[rot, tx, ty]");
    DTransformation::new(f(&a[0]), (f(&a[1]), f(&a[2])))
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let case: Value = serde_json::from_reader(std::fs::File::open(&args[1]).expect("case file")).expect("json");
    let repo_root = args.get(2).map(String::as_str).unwrap_or(".");

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
    let ext = lbf::io::read_spp_instance(&Path::new(repo_root).join(case["instance"].as_str().unwrap())).expect("instance");
    let importer = Importer::new(cde_config, None, None, None);
    let instance = spp::io::import_instance(&importer, &ext).expect("import");

    // the layout: strip of the exported width, placements in order, then the removals
    let mut problem = SPProblem::new(instance.clone());
    problem.change_strip_width(f(&case["strip_width"]));
    let mut pkeys = vec![];
    for p in case["placements"].as_array().unwrap() {
        let pk = problem.place_item(SPPlacement {
            item_id: p["item_id"].as_u64().unwrap() as usize,
            d_transf: dt(&p["d_transf"]),
        });
        pkeys.push(pk);
    }
    for r in case["removed"].as_array().unwrap() {
        problem.remove_item(pkeys[r.as_u64().unwrap() as usize]);
    }
    let cde = problem.layout.cde();

    let (mut n, mut bad_detect, mut bad_collect) = (0usize, 0usize, 0usize);
    for q in case["queries"].as_array().unwrap() {
        let item = instance.item(q["item_id"].as_u64().unwrap() as usize);
        let surrogate = item.shape_cd.surrogate();
        let mut buffer = item.shape_cd.as_ref().clone();
        let exp_d = q["expect_detect"].as_array().unwrap();
        let exp_c = q["expect_collect"].as_array().unwrap();
        for (k, d) in q["d_transfs"].as_array().unwrap().iter().enumerate() {
            let transf = dt(d).compose();
            // detect: lbf/benches/ci_bench.rs:107-121 (bit0 surrogate, bit1 poly - evaluated only when bit0 == 0)
            let mut verdict = 0u64;
            if cde.detect_surrogate_collision(surrogate, &transf, &NoFilter) {
                verdict |= 1;
            } else {
                buffer.transform_from(&item.shape_cd, &transf);
                if cde.detect_poly_collision(&buffer, &NoFilter) {
                    verdict |= 2;
                }
            }
            if verdict != exp_d[k].as_u64().unwrap() {
                bad_detect += 1;
                eprintln!("detect mismatch: item {} query {k}: reference {verdict}, case {}", item.id, exp_d[k]);
            }
            // collect: lbf/benches/ci_bench.rs:56-70 (surrogate and polygon pass share one collector)
            let mut collector = BasicHazardCollector::with_capacity(cde.hazards_map.len());
            cde.collect_surrogate_collisions(surrogate, &transf, &mut collector);
            buffer.transform_from(&item.shape_cd, &transf);
            cde.collect_poly_collisions(&buffer, &mut collector);
            let got: BTreeSet<i64> = collector
                .iter()
                .map(|(_, e)| match e {
                    HazardEntity::Exterior => -1,
                    HazardEntity::PlacedItem { pk, .. } => pkeys.iter().position(|x| x == pk).unwrap() as i64,
                    _ => unreachable!("strip layouts have no holes or zones"),
                })
                .collect();
            let want: BTreeSet<i64> = exp_c[k].as_array().unwrap().iter().map(|v| v.as_i64().unwrap()).collect();
            if got != want {
                bad_collect += 1;
                eprintln!("collect mismatch: item {} query {k}: reference {got:?}, case {want:?}", item.id);
            }
            n += 1;
        }
    }
    println!("{n} queries replayed against jagua-rs: {bad_detect} verdict mismatches, {bad_collect} hazard-set mismatches");
    std::process::exit(if bad_detect + bad_collect == 0 { 0 } else { 1 });
}
