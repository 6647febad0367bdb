// build.rs -- compiles the CUDA library for sm_100a with nvcc and links it.
// (Unbuilt here: no Rust toolchain in the image.  Mirrors the top-level Makefile's `lib` target.)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../csrc");
    let sources = ["rt_kernels.cu", "rt_tile.cu", "rt_phased.cu", "rt_api.cpp", "rt_scene.cpp"];
    let lib = out.join("librtrace_b200.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
        // no FMA contraction on the parity-critical path (the reference is unfused f32)
        .args(["-fmad=false", "-prec-div=true", "-prec-sqrt=true"])
        .args(["-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-x", "cu"])
        .args(sources.iter().map(|s| csrc.join(s)))
        .arg("-o")
        .arg(&lib)
        .status()
        .expect("nvcc not found: the GPU path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rtrace_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    for s in sources.iter() {
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
}
