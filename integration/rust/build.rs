// build.rs for the Wafer crate with the B200 hot path.  UNTESTED here (no Rust toolchain in the build image).
// Original content (vergen, build.rs:1-13 of the reference) is kept; the nvcc step is appended.
extern crate vergen;

use std::env;
use std::process::Command;
use vergen::vergen;

fn main() {
    let mut flags = vergen::OutputFns::all();
    flags.toggle(vergen::COMMIT_DATE);
    flags.toggle(vergen::NOW);
    flags.toggle(vergen::SEMVER);
    flags.toggle(vergen::SHORT_NOW);
    flags.toggle(vergen::TARGET);
    assert!(vergen(flags).is_ok());

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
