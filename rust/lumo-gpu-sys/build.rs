//! Builds liblumo_gpu.so (nvcc, sm_100a) and liblumo_host (cc) from the sources of this repository and links them.
//! The flags are the ones lumo_b200/build.py uses: -fmad=false / -ffp-contract=off are part of the contract (Rust never
//! contracts a * b + c, and the hit ids, distances and films are compared bit for bit with the CPU path).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let so = out.join("liblumo_gpu.so");
    let status = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
               "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I"])
        .arg(root.join("include"))
        .arg("-o").arg(&so)
        .arg(root.join("lumo_b200/csrc/gpu/lumo_gpu.cu"))
        .status().expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    cc::Build::new().cpp(true).flag("-std=c++17").flag("-O2").flag("-ffp-contract=off").flag("-fno-fast-math")
        .file(root.join("lumo_b200/csrc/host/host_build.cpp")).compile("lumo_host");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lumo_gpu");
    println!("cargo:rerun-if-changed={}", root.join("include/lumo_gpu.h").display());
    println!("cargo:rerun-if-changed={}", root.join("lumo_b200/csrc").display());
}
