//! Compiles the CUDA engine with nvcc for sm_100a - the same sources and the same flags as jagua-rs_b200/build.py
//! (bit-exactness needs them: no FMA contraction on either side, IEEE division / square root, no flush-to-zero) - and links
//! it statically. JAGUA_CUDA_SRC points at this repository's `jagua-rs_b200/csrc` (default: ../../jagua-rs_b200/csrc).
//! NOT COMPILED IN THE BUILD IMAGE OF THIS REPOSITORY (no Rust toolchain there).
use std::{env, path::PathBuf, process::Command};

const SOURCES: [&str; 8] = ["kernels.cu", "detect.cu", "search.cu", "engine.cu", "pool.cu", "cde_host.cpp", "cde_image.cpp", "item_host.cpp"];
const HEADERS: [&str; 9] = ["kernels.cuh", "detect_grid.cuh", "collect_grid.cuh", "launch.hpp", "predicates.cuh", "layout_soa.h",
                            "cde_host.hpp", "engine_comm.inl", "../../include/jagua_cuda.h"];

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("JAGUA_CUDA_SRC").unwrap_or_else(|_| "../../jagua-rs_b200/csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                 "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                 "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math"];
    let mut objs = Vec::new();
    for s in SOURCES {
        let obj = out.join(format!("{s}.o"));
        let st = Command::new(&nvcc).args(flags).arg("-c").arg(src.join(s)).arg("-o").arg(&obj)
            .status().expect("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)");
        assert!(st.success(), "nvcc failed on {s}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", src.join(s).display());
    }
    for h in HEADERS {
        println!("cargo:rerun-if-changed={}", src.join(h).display());
    }
    let lib = out.join("libjagua_cuda.a");
    let _ = std::fs::remove_file(&lib);
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().expect("ar");
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=jagua_cuda");
    if let Ok(cuda) = env::var("CUDA_HOME") {
        println!("cargo:rustc-link-search=native={cuda}/lib64");
    } else {
        println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    }
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl"); // NCCL is dlopen'ed at run time (jg_engine_comm_init), never linked
    println!("cargo:rustc-link-lib=dylib=pthread");
}
