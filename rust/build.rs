//! build.rs for the reference crate once `src/graph_b200.rs` is added.
//!
//! Compiles the hand-written CUDA engine for sm_100a with nvcc into a static library and links it,
//! plus the CUDA runtime.  `B200_SPGEMM_DIR` points at this repository's checkout (default: a
//! `b200-spgemm/` directory next to Cargo.toml).  NOT executed in the build image (no cargo there);
//! `sparse_linear_algebra_tests_b200/csrc/Makefile` runs the same nvcc command line.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = env::var("B200_SPGEMM_DIR").map(PathBuf::from).unwrap_or_else(|_| manifest.join("b200-spgemm"));
    let csrc = root.join("sparse_linear_algebra_tests_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".to_string());
    let obj = out.join("api.o");
    let lib = out.join("libb200spgemm.a");

    let status = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-c"])
        .arg(csrc.join("api.cu"))
        .arg("-o")
        .arg(&obj)
        .status()
        .expect("nvcc not found: set NVCC or install CUDA 12.8+ (sm_100a)");
    assert!(status.success(), "nvcc failed on api.cu");
    let status = Command::new("ar").arg("rcs").arg(&lib).arg(&obj).status().expect("ar not found");
    assert!(status.success(), "ar failed");

    let cuda_lib = env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".to_string());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={cuda_lib}");
    println!("cargo:rustc-link-lib=static=b200spgemm");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for f in ["api.cu", "kernels.cuh", "common.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("b200_spgemm.h").display());
    println!("cargo:rerun-if-env-changed=B200_SPGEMM_DIR");
    println!("cargo:rerun-if-env-changed=NVCC");
}
