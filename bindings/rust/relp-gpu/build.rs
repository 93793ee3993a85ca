// Links librelp_gpu.so (built by `python -m relp_b200.build`); RELP_GPU_LIB_DIR points at relp_b200/.
fn main() {
    let dir = std::env::var("RELP_GPU_LIB_DIR").unwrap_or_else(|_| "../../../relp_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=relp_gpu");
    println!("cargo:rerun-if-env-changed=RELP_GPU_LIB_DIR");
    #[cfg(feature = "bindgen")]
    {
        let header = "../../../include/relp_gpu.h";
        println!("cargo:rerun-if-changed={header}");
        bindgen::Builder::default()
            .header(header)
            .allowlist_function("rg_.*")
            .allowlist_type("rg_.*")
            .allowlist_var("RG_.*")
            .generate()
            .expect("bindgen over include/relp_gpu.h")
            .write_to_file("src/ffi.rs")
            .expect("write src/ffi.rs");
    }
}
