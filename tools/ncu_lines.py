#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output: per CUDA source line,
warp instructions executed and stall samples, for kernels matching a substring.
usage: ncu_lines.py <src.csv> <kernel-substring> [top_n]"""
import csv
import sys

path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path, errors="replace")))
i, seen = 0, 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Function Name" and pat in r[1]:
        seen += 1
        print("#### file:", rows[i - 1][1] if i > 0 and rows[i - 1] else "?")
        hdr = rows[i + 1]
        col = {n: k for k, n in enumerate(hdr)}
        ci, cs = col["Instructions Executed"], col["# Samples"]
        stall_cols = [k for k, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        j = i + 2
        lines, tot_i, tot_s = [], 0, 0
        stalls_tot = {}
        while j < len(rows) and rows[j] and rows[j][0] not in ("File Path", "Function Name", "File Name"):
            q = rows[j]
            if q[0] != "" and len(q) > ci:  # a CUDA source line aggregate row
                try:
                    n_i, n_s = int(q[ci]), int(q[cs])
                except ValueError:
                    j += 1
                    continue
                st = {hdr[k][6:]: int(q[k]) for k in stall_cols if q[k] not in ("", "0")}
                lines.append((n_i, n_s, q[0], q[1].strip()[:90], st))
                tot_i += n_i
                tot_s += n_s
                for k, v in st.items():
                    stalls_tot[k] = stalls_tot.get(k, 0) + v
            j += 1
        print(f"== {r[1][:80]}  total warp-instr {tot_i}  samples {tot_s}")
        print("   stalls:", sorted(stalls_tot.items(), key=lambda x: -x[1])[:8])
        for n_i, n_s, ln, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
            s3 = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
            print(f"   {100*n_i/max(tot_i,1):5.1f}%i {100*n_s/max(tot_s,1):5.1f}%s L{ln:>4} {src}  [{s3}]")
        i = j
    else:
        i += 1
