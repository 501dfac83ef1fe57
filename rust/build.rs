// build.rs for the reference crate: builds libcity2ba_cuda.so with nvcc and links it (replaces the
// Embree link lines, reference build.rs:4-11).  NOT compiled in this image (no rustc/cargo).
use std::{env, path::PathBuf, process::Command};

fn main() {
    // CITY2BA_CUDA_DIR = checkout of this repository
    let root = PathBuf::from(env::var("CITY2BA_CUDA_DIR").expect("set CITY2BA_CUDA_DIR"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libcity2ba_cuda.so");
    let status = Command::new("nvcc")
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o"])
        .arg(&so)
        .arg(root.join("city2ba_b200/csrc/c2b_api.cu"))
        .arg(root.join("city2ba_b200/csrc/c2b_multi.cu"))
        .arg(root.join("city2ba_b200/csrc/c2b_host.cpp"))
        .arg("-ldl")  // libnccl.so.2 is loaded with dlopen by c2b_init_multi only
        .status().expect("nvcc not found");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=city2ba_cuda");
    println!("cargo:rerun-if-changed={}", root.join("city2ba_b200/csrc").display());
}
