// Links libqgpu.so (built by `python -m query_engine_b200.build`, plain nvcc for sm_100a).
// QGPU_LIB_DIR points at the directory that holds it; default: ../../query_engine_b200/lib relative to this crate.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("QGPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../query_engine_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=qgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=QGPU_LIB_DIR");
}
