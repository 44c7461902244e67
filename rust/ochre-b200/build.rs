// build.rs -- compiles the CUDA sources of the repository with nvcc for sm_100a and links the
// resulting shared library.  NOTE: written but never compiled in the build image (no rustc/cargo
// there); it performs exactly the command ochre_b200/build.py runs.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("ochre_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libochre_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let status = Command::new(nvcc)
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
                "--extended-lambda", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-o"])
        .arg(&so)
        .arg(csrc.join("pipeline.cu"))
        .arg(csrc.join("host_path.cpp"))
        .status()
        .expect("nvcc not found: ochre-b200 has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ochre_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
