// Replaces the reference's build.rs step that spawns $REST_FORTRAN_COMPILER on restmatr.f90 (reference build.rs:25-39):
// build librest_b200.so with nvcc for sm_100a and link it (plus cudart, which nvcc links statically into the .so).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let repo = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = repo.join("rest_tensors_b200/csrc");
    let out_dir = env::var("REST_EXT_DIR").map(PathBuf::from).unwrap_or_else(|_| repo.join("rest_tensors_b200"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| {
        format!("{}/bin/nvcc", env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into()))
    });
    let status = Command::new("make")
        .arg("-C").arg(&csrc).arg("-j8").arg(format!("NVCC={}", nvcc))
        .status().expect("failed to run make for librest_b200.so");
    assert!(status.success(), "nvcc build of librest_b200.so failed");
    println!("cargo:rustc-link-search=native={}", out_dir.display());
    println!("cargo:rustc-link-lib=dylib=rest_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", repo.join("include/rest_b200.h").display());
}
