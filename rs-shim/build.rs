// rs/build.rs - link the B200 backend.  SWFR_B200_LIB_DIR = directory holding libswfr_b200.so
// (swf_renderer_b200/ in the backend's repository after `make -C swf_renderer_b200/csrc`).
fn main() {
  let dir = std::env::var("SWFR_B200_LIB_DIR").unwrap_or_else(|_| "../swf_renderer_b200".to_string());
  println!("cargo:rustc-link-search=native={}", dir);
  println!("cargo:rustc-link-lib=dylib=swfr_b200");
  println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
  println!("cargo:rerun-if-env-changed=SWFR_B200_LIB_DIR");
}
