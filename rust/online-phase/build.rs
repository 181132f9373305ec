//! Links `libarkmpc_b200.so` (built by `python -c "import __graft_entry__ as g; g.build()"`: nvcc, sm_100a only) when the `b200`
//! feature is on.  Same role as mp-spdz-rs/build.rs:15-63 for the MP-SPDZ bridge.
fn main() {
    println!("cargo:rerun-if-env-changed=ARKMPC_B200_LIB_DIR");
    if std::env::var("CARGO_FEATURE_B200").is_err() {
        return;
    }
    let dir = std::env::var("ARKMPC_B200_LIB_DIR").expect("feature `b200`: set ARKMPC_B200_LIB_DIR to <ark-mpc-b200>/ark_mpc_b200/lib");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=arkmpc_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
