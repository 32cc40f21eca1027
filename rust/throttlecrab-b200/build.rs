// NOT compiled in this repository's environment (no rustc/cargo in the image): source for maintainers.
fn main() {
    // libgcra_b200.so is built by nvcc (see __graft_entry__.build()); just link it.
    println!("cargo:rustc-link-search=native={}", std::env::var("GCRA_B200_LIB_DIR").unwrap());
    println!("cargo:rustc-link-lib=dylib=gcra_b200");
    // bindgen alternative: bindgen::Builder::default().header("include/gcra_b200.h")...
}
