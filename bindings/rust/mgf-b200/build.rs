// Link against the library this repository builds (python -c "import __graft_entry__ as g; g.build()").
fn main() {
    let dir = std::env::var("MGFB_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../../mgf_b200/lib").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=mgfb");
    println!("cargo:rerun-if-env-changed=MGFB_LIB_DIR");
}
