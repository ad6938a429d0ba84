// build.rs -- compiles the CUDA library for sm_100a with nvcc and links it statically.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the development image has no Rust toolchain.  The same
// objects are built by rust-modem_b200/csrc/Makefile and exercised through the identical
// extern "C" symbols by the C++ mirror (rust-modem_b200/host) and the Python tests;
// tests/test_rust_boundary.py checks that this file compiles every unit the Makefile does (the unit
// list is not written down here: every csrc/*.cu is a unit) and with the same tuning defines.
//
// Drop this file next to the reference's Cargo.toml and add `build = "build.rs"` to [package].
use std::env;
use std::path::PathBuf;
use std::process::Command;

// keep in step with RXTUNE in csrc/Makefile (checked by tests/test_rust_boundary.py)
const RXTUNE: [&str; 6] = ["-DRX_DEFAULT_THREADS=64", "-DRX_DEFAULT_MINB=8", "-DRX_DEFAULT_R=4", "-DRX_DEFAULT_PF=5", "-DRX_CARVEOUT=66", "-DRX_DEFAULT_TMC=64"];

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("MODEM_GPU_CSRC").unwrap_or_else(|_| "rust-modem_b200/csrc".into()));
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    let nvcc = format!("{}/bin/nvcc", cuda);
    // every *.cu under csrc/ is one translation unit of the library (the Makefile's CUOBJS)
    let mut units: Vec<PathBuf> = std::fs::read_dir(&csrc)
        .expect("csrc directory not found (set MODEM_GPU_CSRC)")
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    units.sort();
    assert!(!units.is_empty(), "no .cu units under {}", csrc.display());
    let mut objs = Vec::new();
    for u in units.iter() {
        let o = out.join(u.file_stem().unwrap()).with_extension("o");
        let st = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false"])
            .args(["-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"])
            .args(RXTUNE)
            .arg("-c").arg("-o").arg(&o).arg(u)
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {}", u.display());
        objs.push(o);
    }
    // host-side tables: the reference's mapper formulas, unfused binary32
    let ht = out.join("host_tables.o");
    let st = Command::new("g++")
        .args(["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-c", "-o"]).arg(&ht)
        .arg(csrc.join("host_tables.cpp")).status().expect("g++ not found");
    assert!(st.success());
    objs.push(ht);
    let lib = out.join("libmodem_gpu.a");
    let _ = std::fs::remove_file(&lib);
    let st = Command::new("ar").arg("rcs").arg(&lib).args(&objs).status().unwrap();
    assert!(st.success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=modem_gpu");
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
    // NCCL is dlopen()ed by the library at the first modem_gpu_comm_* call: no link-time dependency.
    println!("cargo:rerun-if-changed={}", csrc.display());
}
