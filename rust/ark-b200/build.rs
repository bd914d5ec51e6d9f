fn main() {
    // B200SNARK_LIB_DIR = directory holding libb200snark.so (snark_b200/ in this repository)
    let dir = std::env::var("B200SNARK_LIB_DIR").expect("set B200SNARK_LIB_DIR");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b200snark");
}
