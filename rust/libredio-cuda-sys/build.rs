// Tell cargo where libredio_cuda.so lives (built by `python -m libredio_b200.build`).
fn main() {
    let dir = std::env::var("LIBREDIO_CUDA_DIR").unwrap_or_else(|_| "../../libredio_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=redio_cuda");
    println!("cargo:rerun-if-env-changed=LIBREDIO_CUDA_DIR");
}
