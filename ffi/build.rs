// build.rs -- compiles the CUDA sources with nvcc for sm_100a and links the resulting shared
// library.  Replaces /root/reference/build.rs (cmake crate -> damavand-gpu/CMakeLists.txt, which
// passes no -arch flag at all).  Same command line as damavand_b200/build.py.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("..").join("damavand_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libdamavand_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let sources = ["kernels.cu", "engine.cu", "compat.cu", "planner.cpp"];
    let mut cmd = Command::new(nvcc);
    cmd.args(&[
        "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC,-O3", "-shared", "-cudart", "static", "-o",
    ]);
    cmd.arg(&lib);
    for s in &sources {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    cmd.arg("-ldl");
    let status = cmd.status().expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=damavand_b200");
}
