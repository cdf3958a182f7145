// Links liblabrador_ldpc.so (built by `python -m labrador_ldpc_b200._build`).
fn main() {
    let dir = std::env::var("LABRADOR_LDPC_B200_LIB_DIR")
        .unwrap_or_else(|_| "../../labrador_ldpc_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=labrador_ldpc");
}
