// SOURCE ONLY (never compiled here: no Rust toolchain in the image).
// Builds libheic_b200.so with the same nvcc line __graft_entry__.build() uses and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut sources = vec![];
    for dir in ["heif_b200/csrc/host", "heif_b200/csrc/cuda"] {
        for e in std::fs::read_dir(root.join(dir)).unwrap() {
            let p = e.unwrap().path();
            match p.extension().and_then(|s| s.to_str()) {
                Some("cc") | Some("cu") => sources.push(p),
                _ => {}
            }
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
    let lib = out.join("libheic_b200.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-x", "cu"])
        .arg(format!("-I{}", root.join("include").display()))
        .args(&sources)
        .arg("-o")
        .arg(&lib)
        .status()
        .expect("nvcc not found: the B200 path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=heic_b200");
}
