//! Compiles the CUDA sources for sm_100a with nvcc and links the resulting shared library.
//! Equivalent to `python -m lair_b200.build` (lair_b200/build.py), which is what this
//! repository's own tests use because no Rust toolchain is available in its build image.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("lair_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".to_string());
    let mut objects = Vec::new();
    for entry in std::fs::read_dir(&csrc).expect("csrc directory") {
        let path = entry.unwrap().path();
        if path.extension().map(|e| e == "cu").unwrap_or(false) {
            let obj = out.join(path.file_stem().unwrap()).with_extension("o");
            let status = Command::new(&nvcc)
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
                .args(["-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"])
                .arg("-I")
                .arg(root.join("include"))
                .arg("-c")
                .arg(&path)
                .arg("-o")
                .arg(&obj)
                .status()
                .expect("failed to run nvcc");
            assert!(status.success(), "nvcc failed for {}", path.display());
            println!("cargo:rerun-if-changed={}", path.display());
            objects.push(obj);
        }
    }
    let lib = out.join("liblair_b200.so");
    let status = Command::new(&nvcc)
        .args(["-shared", "-o"])
        .arg(&lib)
        .args(&objects)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"])
        .status()
        .expect("failed to link");
    assert!(status.success(), "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lair_b200");
}
