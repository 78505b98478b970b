// build.rs for the Wafer crate with the B200 hot path.  UNTESTED here (no Rust toolchain in the build image).
// The reference's own content (the vergen block) is not repeated here; the nvcc step below is appended to it.
use std::env;
use std::process::Command;

fn main() {
    // (1) keep the crate's existing version-stamp block here unchanged (build.rs:1-13 of the reference: the
    //     `vergen` call that generates version.rs for main.rs:66,200).

    // (2) new: build the CUDA library.
    // sm_100a only: there is no CPU fallback and no other architecture in the fat binary
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    let out = env::var("OUT_DIR").unwrap();
    let lib = format!("{}/libwafer_b200.so", out);
    let status = Command::new(format!("{}/bin/nvcc", cuda))
        .args(&[
            "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
            "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", "-o", &lib,
            "wafer_b200/csrc/wafer_b200.cu", "-ldl",
        ])
        .status()
        .expect("nvcc not found: set CUDA_HOME");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out);
    println!("cargo:rustc-link-lib=dylib=wafer_b200"); // cudart is linked statically into the .so; NCCL is dlopen'ed
    println!("cargo:rerun-if-changed=wafer_b200/csrc");
    println!("cargo:rerun-if-changed=include/wafer_b200.h");
}
