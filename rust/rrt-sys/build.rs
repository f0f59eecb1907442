// Links librrt_sm100.so.  RRT_LIB_DIR = the directory that holds it (rs_ray_toy_b200/ in the source tree).
fn main() {
    if let Ok(dir) = std::env::var("RRT_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=rrt_sm100");
    println!("cargo:rerun-if-env-changed=RRT_LIB_DIR");
}
