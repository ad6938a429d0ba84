"""Guards of the Rust side of the boundary (rust-modem_b200/rust/), which this image cannot compile (no rustc):
  (i)  build.rs compiles the translation units the Makefile compiles, with the same tuning defines;
  (ii) src/gpu.rs binds every function include/modem_gpu.h declares -- same name, same number of arguments, pointer
       arguments where the header has pointers -- and mirrors modem_cfg_t / modem_phasor_t field for field.
Round 1 shipped a build.rs whose unit list missed two units (an archive that could not link) and a gpu.rs with 24 of
50 symbols; these tests are what would have caught both."""
import os
import re

from conftest import ROOT

RUST = os.path.join(ROOT, "rust-modem_b200", "rust")
CSRC = os.path.join(ROOT, "rust-modem_b200", "csrc")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def _split_args(arglist):
    out, depth, cur = [], 0, ""
    for ch in arglist:
        if ch in "([<":
            depth += 1
        elif ch in ")]>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_functions():
    src = _strip_c_comments(open(os.path.join(ROOT, "include", "modem_gpu.h")).read())
    fns = {}
    for m in re.finditer(r"\b(modem_(?:gpu_)?[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        params = [] if args in ("", "void") else _split_args(args)
        fns[name] = ["*" in p or "[" in p for p in params]
    return fns


def rust_externs():
    src = open(os.path.join(RUST, "src", "gpu.rs")).read()
    src = re.sub(r"//.*", "", src)
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    fns = {}
    for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        params = _split_args(" ".join(m.group(2).split()))
        fns[m.group(1)] = ["*" in p.split(":", 1)[1] for p in params]
    return fns


def test_gpu_rs_binds_every_header_function():
    h, r = header_functions(), rust_externs()
    assert len(h) >= 52, sorted(h)
    assert sorted(set(h) - set(r)) == [], "declared in modem_gpu.h but not bound in gpu.rs"
    assert sorted(set(r) - set(h)) == [], "bound in gpu.rs but not declared in modem_gpu.h"
    for name in h:
        assert len(h[name]) == len(r[name]), (name, "argument count", len(h[name]), len(r[name]))
        assert h[name] == r[name], (name, "pointer / value arguments differ", h[name], r[name])


def _struct_fields_c(name):
    src = _strip_c_comments(open(os.path.join(ROOT, "include", "modem_gpu.h")).read())
    body = re.search(r"typedef struct \{([^}]*)\}\s*" + name + r"\s*;", src, flags=re.S).group(1)
    fields = []
    for decl in filter(None, (d.strip() for d in body.split(";"))):
        parts = [p.strip() for p in decl.split(",")]  # "uint8_t start, end" declares two fields
        for p in parts:
            fields.append((p.split()[-1].lstrip("*"), "*" in decl))
    return fields


def _struct_fields_rust(name):
    src = re.sub(r"//.*", "", open(os.path.join(RUST, "src", "gpu.rs")).read())
    body = re.search(r"pub struct " + name + r" \{(.*?)\n\}", src, flags=re.S).group(1)
    return [(m.group(1), "*" in m.group(2)) for m in re.finditer(r"pub (\w+):\s*([^,\n]+)", body)]


def test_gpu_rs_structs_mirror_the_header():
    for name in ("modem_cfg_t", "modem_phasor_t", "modem_ring_t", "modem_c32_t"):
        assert _struct_fields_c(name) == _struct_fields_rust(name), name


def test_gpu_rs_constants_match_the_header():
    h = open(os.path.join(ROOT, "include", "modem_gpu.h")).read()
    r = open(os.path.join(RUST, "src", "gpu.rs")).read()
    for flag in ("MODEM_FLAG_FUSED_MAC", "MODEM_FLAG_NO_TMEM"):
        hv = int(re.search(r"#define " + flag + r"\s+(0x[0-9a-fA-F]+)u", h).group(1), 16)
        rv = int(re.search(r"pub const " + flag + r": u32 = (0x[0-9a-fA-F]+);", r).group(1), 16)
        assert hv == rv, flag
    assert int(re.search(r"#define MODEM_COMM_ID_BYTES (\d+)", h).group(1)) == int(re.search(r"MODEM_COMM_ID_BYTES: usize = (\d+)", r).group(1))
    for i, fmt in enumerate(("MODEM_SAMPLES_C32", "MODEM_SAMPLES_F32", "MODEM_SAMPLES_I16")):
        assert re.search(fmt + r" = " + str(i), h) and re.search(r"pub const " + fmt + r": u32 = " + str(i), r)


def test_build_rs_compiles_what_the_makefile_compiles():
    mk = open(os.path.join(CSRC, "Makefile")).read()
    rs = open(os.path.join(RUST, "build.rs")).read()
    # both take EVERY csrc/*.cu as a unit (no list to go stale) ...
    assert "$(wildcard *.cu)" in mk and "CUOBJS" in mk
    assert "read_dir" in rs and 'x == "cu"' in rs
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    assert {"modem_api.cu", "tx_fast.cu", "rx_fast_64.cu", "rx_fast_129.cu", "loop_fused_64.cu", "rx_fullrate_fast.cu", "rx_fast_dispatch.cu"} <= set(units)
    # ... every unit defines something the API unit needs or is the API unit: no stray sources that would not link
    for u in units:
        assert 'namespace mg' in open(os.path.join(CSRC, u)).read() or u == "modem_api.cu", u
    # ... with the same tuning defines and code-generation flags
    tune_mk = sorted(re.findall(r"-DRX_\w+=\d+", re.search(r"RXTUNE\s*\?=\s*(.*)", mk).group(1)))
    tune_rs = sorted(re.findall(r"-DRX_\w+=\d+", rs))
    assert tune_mk == tune_rs and len(tune_mk) == 6
    for flag in ("arch=compute_100a,code=sm_100a", "-fmad=false", "-ffp-contract=off", "-std=c++17", "-lineinfo"):
        assert flag in mk and flag in rs, flag
    assert "host_tables.cpp" in mk and "host_tables.cpp" in rs


def test_header_cites_the_reference_for_every_path_entry():
    """include/modem_gpu.h must say which reference lines each compute entry replaces (the drop-in contract)."""
    src = open(os.path.join(ROOT, "include", "modem_gpu.h")).read()
    for entry, cite in (("modem_gpu_modulate(", "modulator.rs"), ("modem_gpu_demodulate(", "demodulator.rs"), ("modem_gpu_lock_phase(", "pll.rs"),
                        ("modem_gpu_preamble(", "modulate.rs:118-126"), ("modem_gpu_modulate_real(", "modulate.rs:118-133"),
                        ("modem_gpu_demodulate_real(", "demodulate.rs:29-43")):
        i = src.index(entry)
        assert cite in src[max(0, i - 1600):i], (entry, cite)
