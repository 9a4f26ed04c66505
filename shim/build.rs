// links libzkb200.so (built by `make -C ziren_b200/csrc`); ZKB200_LIB_DIR points at the directory holding it
fn main() {
    if let Ok(dir) = std::env::var("ZKB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
