// Links against the in-tree build of the CUDA library (co_snarks_b200/libcosnarks_gpu.so).
fn main() {
    let dir = std::env::var("COSNARKS_GPU_LIB_DIR").unwrap_or_else(|_| "../../co_snarks_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cosnarks_gpu");
    println!("cargo:rerun-if-env-changed=COSNARKS_GPU_LIB_DIR");
}
