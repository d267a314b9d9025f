// build.rs — compiles the CUDA sources of frave_b200/csrc into libfri_cuda.a with nvcc for
// sm_100a and links it (the same flags frave_b200/build.py uses).  Untested here: no cargo.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("frave_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libfri_cuda.a");
    let status = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC,-ffp-contract=off", "--lib", "-o"])
        .arg(&lib)
        .arg(csrc.join("fri_api.cu"))
        .arg(csrc.join("fri_kernels.cu"))
        .arg(csrc.join("fri_predict.cu"))
        .arg(csrc.join("fri_plan.cpp"))
        .arg(csrc.join("fri_order.cpp"))
        .arg(csrc.join("fri_codec.cpp"))
        .status()
        .expect("nvcc not found: libfri-cuda has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=fri_cuda");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for f in ["fri_api.cu", "fri_kernels.cu", "fri_predict.cu", "fri_plan.cpp", "fri_order.cpp", "fri_codec.cpp", "fri_codec.h", "fri_kernels.cuh", "fri_plan.h", "fri_geometry.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
}
