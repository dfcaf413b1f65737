// libtagg.so is built by build.sh (nvcc, sm_100a) into tantivy_aggregations_b200/; point TAGG_LIB_DIR at it.
fn main() {
    if let Ok(dir) = std::env::var("TAGG_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=tagg");
    println!("cargo:rerun-if-env-changed=TAGG_LIB_DIR");
}
