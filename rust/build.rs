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
    let lib = out.join("libb200spgemm.a");

    // one object per translation unit, as in csrc/Makefile
    let units = ["api", "fused", "rowwarp", "heavy", "coo", "comm", "dense", "leftmul"];
    let mut objs = Vec::new();
    for u in units {
        let obj = out.join(format!("{u}.o"));
        let status = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(format!("{u}.cu")))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: set NVCC or install CUDA 12.8+ (sm_100a)");
        assert!(status.success(), "nvcc failed on {u}.cu");
        objs.push(obj);
    }
    let status = Command::new("ar").arg("rcs").arg(&lib).args(&objs).status().expect("ar not found");
    assert!(status.success(), "ar failed");

    let cuda_lib = env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".to_string());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={cuda_lib}");
    println!("cargo:rustc-link-lib=static=b200spgemm");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");       // comm.cu binds NCCL at run time (dlopen of libnccl.so.2)
    for f in ["api.cu", "fused.cu", "rowwarp.cu", "heavy.cu", "coo.cu", "comm.cu", "dense.cu", "leftmul.cu", "kernels.cuh", "devutil.cuh", "engine.cuh", "gen.cuh", "common.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("b200_spgemm.h").display());
    println!("cargo:rerun-if-env-changed=B200_SPGEMM_DIR");
    println!("cargo:rerun-if-env-changed=NVCC");
}
