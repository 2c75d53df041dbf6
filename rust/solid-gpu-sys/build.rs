// Compiles the CUDA sources of the filtering hot path for sm_100a and links them statically.
// Requires nvcc (CUDA >= 12.8) on PATH or under $CUDA_HOME.  There is no CPU fallback to build.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../solid_dsp_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let cuda_home = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    let nvcc = PathBuf::from(&cuda_home).join("bin/nvcc");
    // keep in step with SRCS in solid_dsp_b200/csrc/Makefile
    let sources = ["common.cu", "fir.cu", "fir_tc.cu", "iir.cu", "dot.cu", "autocorr.cu", "nco.cu", "ctx.cu", "firdes.cu"];
    let mut objects = Vec::new();
    for src in sources.iter() {
        let obj = out.join(format!("{}.o", src));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("failed to run nvcc");
        assert!(status.success(), "nvcc failed on {}", src);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objects.push(obj);
    }
    let lib = out.join("libsolid_gpu.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().expect("ar");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=solid_gpu");
    println!("cargo:rustc-link-search=native={}/lib64", cuda_home);
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", manifest.join("../../include/solid_gpu.h").display());
}
