#!/usr/bin/env python
"""Summarise registers / spills per kernel from rust-modem_b200/lib/*.ptxas.log"""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rust-modem_b200", "lib")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(log).read()
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \1\n\s+(.*)\nptxas info\s+: Used (\d+) registers(.*)", txt)
    for name, props, regs, rest in ents:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void mg::", "")
        if pat in dem:
            sp = re.search(r"(\d+) bytes spill stores", props).group(1)
            print(f"{dem:66s} regs={regs:>3s} spill={sp:>4s} {rest.strip()[:40]}")
