// build.rs — compiles whisper-diarize-rs_b200/csrc/*.cu for sm_100a (nvcc, via the csrc Makefile) and links libwdr_b200.a.
// The reference crate has no build script of its own: whisper-rs-sys / knf-rs-sys / ort-sys each ran one (cmake, prebuilt
// ONNX Runtime download); this single script replaces all three.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../whisper-diarize-rs_b200/csrc").canonicalize().expect("csrc directory");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new("make")
        .arg("-C").arg(&csrc).arg("libwdr_b200.a").arg(format!("NVCC={nvcc}"))
        .status()
        .expect("make / nvcc not found: libwdr_b200 needs the CUDA 12.9 toolkit");
    assert!(status.success(), "nvcc failed (the library targets sm_100a only: -gencode arch=compute_100a,code=sm_100a)");
    std::fs::copy(csrc.join("libwdr_b200.a"), out.join("libwdr_b200.a")).unwrap();
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=wdr_b200");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");  // NCCL is opened at run time (csrc/dist.cu): no link-time dependency on libnccl
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", manifest.join("../../include/wdr.h").display());
}
