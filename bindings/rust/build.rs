// build.rs additions for src-tauri (the reference's build.rs only adds macOS Swift rpaths,
// src-tauri/build.rs:1-18).  Builds libcrispy_ns.so with nvcc for sm_100a and links it.
// NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CRISPY_NS_SRC").expect("CRISPY_NS_SRC = checkout of this repository"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libcrispy_ns.so");
    let csrc = root.join("crispy_b200/csrc");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("crispy_ns.cu"))
        .arg(csrc.join("ns_host.cpp"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=crispy_ns");
    println!("cargo:rerun-if-changed={}", csrc.display());
    tauri_build::build()
}
