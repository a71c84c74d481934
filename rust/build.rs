// build.rs of the `gpu` feature: compiles the .cu sources of libsspsd for sm_100a with nvcc and
// links them.  SOURCE ONLY -- there is no Rust toolchain in the build container, so this file has
// never been compiled; it shows the binding a stabilizer-stream maintainer would add.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("SSPSD_ROOT").unwrap_or_else(|_| "../stabilizer_stream_b200".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libsspsd.a");
    let objs: Vec<PathBuf> = ["sspsd_cascade", "sspsd_api", "sspsd_source", "sspsd_receiver", "sspsd_group"]
        .iter()
        .map(|name| {
            let obj = out.join(format!("{name}.o"));
            let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
                .args(["-Xcompiler", "-fPIC", "-c"])
                .arg(root.join("csrc").join(format!("{name}.cu")))
                .arg("-o")
                .arg(&obj)
                .status()
                .expect("nvcc not found");
            assert!(status.success(), "nvcc failed on {name}.cu");
            obj
        })
        .collect();
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=sspsd");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rustc-link-lib=dl"); // NCCL is dlopen()ed by the first multi-GPU group (libnccl.so.2), never linked
    println!("cargo:rerun-if-changed={}", root.join("csrc").display());
}
