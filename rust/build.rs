// build.rs — replaces the reference's build.rs:1-4 (which only ran tonic-build on the protobufs).
// Adds the native step the reference never had: nvcc compiles the hand-written sm_100a kernels into
// libfd_b200.so and the crate links it.  (The tonic-build lines stay as they were, with the directory name
// fixed: the reference points at `triton-proto/` but ships `triton_proto/`.)
//
// Every *.cu under fd_b200/csrc is compiled (the same set rs_face_detection_b200/build.py lists; tests/test_abi.py
// checks that the two agree), and the link step refuses undefined symbols (-z defs), so a missing translation unit
// fails the BUILD instead of the first call at run time.
use std::{env, fs, path::PathBuf, process::Command};

fn main() -> Result<(), Box<dyn std::error::Error>> {
    tonic_build::compile_protos("triton_proto/grpc_service.proto")?;
    tonic_build::compile_protos("triton_proto/model_config.proto")?;

    let out = PathBuf::from(env::var("OUT_DIR")?);
    let csrc = PathBuf::from("fd_b200/csrc"); // = rs_face_detection_b200/csrc of this repo, vendored into the crate
    let mut sources: Vec<PathBuf> = fs::read_dir(&csrc)?
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    sources.sort();
    assert!(!sources.is_empty(), "no CUDA sources under {}", csrc.display());
    for entry in fs::read_dir(&csrc)? {
        println!("cargo:rerun-if-changed={}", entry?.path().display()); // .cu and the .cuh headers they include
    }
    println!("cargo:rerun-if-changed=fd_b200/include/fd_b200.h");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libfd_b200.so");
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--shared", "-cudart", "shared",
              "-Xlinker", "-z", "-Xlinker", "defs", "-o"])
        .arg(&lib);
    for s in &sources {
        cmd.arg(s);
    }
    assert!(cmd.status()?.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=fd_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    Ok(())
}
