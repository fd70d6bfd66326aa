// Compiles the engine with nvcc for sm_100a and links it (BASELINE.json north_star: "a thin extern "C" FFI compiled by
// build.rs with nvcc").  PTTS_CUDA_SRC points at pocket_tts_b200/csrc of this repository (default: ../../pocket_tts_b200).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("PTTS_CUDA_SRC").unwrap_or_else(|_| "../../pocket_tts_b200".into()));
    let lib = out.join("libptts_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
               "-diag-suppress", "177"])
        .arg(root.join("csrc/engine.cu"))
        .arg("-o")
        .arg(&lib)
        .status()
        .expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptts_cuda");
    println!("cargo:rerun-if-changed={}", root.join("csrc").display());
    println!("cargo:rerun-if-changed=../../include/ptts.h");
}
