#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the evidence that the hot kernels are sm_100a
code using tensor memory (LDTM / STTM), bulk L2 prefetch (UBLKPF), line prefetch into L1 (CCTL.E.PF1), packed binary32 (FMUL2 / FFMA2) and, where the angle
depends on the data, the binary64 libm (DFMA); and that no tensor-core (UTC*MMA / HMMA) or TMA tile copy (UTMALDG) opcode
is used -- the path is a stencil, not a GEMM.   usage: sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rust-modem_b200", "lib", "libmodem_gpu.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["FMUL2", "FFMA2", "FADD2", "FMUL", "FFMA", "FADD", "DFMA", "DMUL", "DADD", "LDTM", "STTM", "UTCBAR", "UBLKPF", "CCTL", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "HMMA", "IMAD",
       "LDG", "STG", "LDS", "STS", "LDC", "LDCU", "SHFL", "BAR", "MUFU", "ATOMG", "RED", "ATOMS"]
funcs, cur, arch = collections.OrderedDict(), None, None
for line in txt.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (arch {arch}); {len(funcs)} kernels; static instruction counts per kernel")
print("# columns: total | " + " ".join(KEY))
tot = collections.Counter()
for f, c in funcs.items():
    tot.update(c)
    name = demangle(f)
    name = (name[:name.rindex(">(") + 1] if ">(" in name else re.sub(r"\(.*", "", name)).replace("void mg::", "").replace("(int)", "").replace("(bool)", "")
    print(f"{name[:96]:96s} {sum(c.values()):6d} | " + " ".join(f"{k}={c[k]}" for k in KEY if c[k]))
print()
print("whole library: " + " ".join(f"{k}={tot[k]}" for k in KEY if tot[k]) + f"  total={sum(tot.values())}")
print("tensor-memory management (tcgen05.alloc / dealloc / relinquish): " + (", ".join(f"{k}={v}" for k, v in tot.items() if k.startswith("UTCATOM")) or "none"))
print("tensor-core MMA / TMA tile-copy opcodes (UTC*MMA, HMMA, UTMALDG, UTMASTG): " + (", ".join(f"{k}={v}" for k, v in tot.items() if k.startswith(("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMA", "HMMA", "QMMA", "OMMA"))) or "none"))
