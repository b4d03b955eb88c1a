// UNTESTED in the build environment of this repository (no Rust toolchain there).
// Links libzkm_b200.so; set ZKM_B200_LIB_DIR to <repo>/zkmember_b200/lib.
fn main() {
    let dir = std::env::var("ZKM_B200_LIB_DIR").expect("set ZKM_B200_LIB_DIR to the directory holding libzkm_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=zkm_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=ZKM_B200_LIB_DIR");
}
