//! Builds (python + nvcc, sm_100a) and links libgs_b200.so.  GS_B200_ROOT = checkout of this repository
//! (defaults to the parent directory of this crate).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = env::var("GS_B200_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join(".."));
    let pkg = root.join("groth-sahai-rs_b200");
    let status = Command::new("python3")
        .arg(pkg.join("build.py")) // nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo ...
        .status()
        .expect("could not run groth-sahai-rs_b200/build.py (python3 + nvcc needed)");
    assert!(status.success(), "building libgs_b200.so failed");
    println!("cargo:rustc-link-search=native={}", pkg.display());
    println!("cargo:rustc-link-lib=dylib=gs_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", pkg.display());
    println!("cargo:rerun-if-changed={}", root.join("include/gs_b200.h").display());
    println!("cargo:rerun-if-changed={}", pkg.join("csrc").display());
}
