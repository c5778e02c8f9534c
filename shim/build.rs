// build.rs -- link libministark.so (built by `python -m ministark_b200.build`: nvcc, sm_100a) when the `b200` feature is on.
// MINISTARK_B200_LIB_DIR = the directory that holds libministark.so (e.g. <this repo>/ministark_b200).
fn main() {
    println!("cargo:rerun-if-env-changed=MINISTARK_B200_LIB_DIR");
    if std::env::var("CARGO_FEATURE_B200").is_ok() {
        let dir = std::env::var("MINISTARK_B200_LIB_DIR").expect("set MINISTARK_B200_LIB_DIR to the directory of libministark.so");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=ministark");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
}
