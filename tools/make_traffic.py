#!/usr/bin/env python
"""profiles/traffic.json from ncu reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of the bench's kernels
(what bench.py reports as roofline.traffic), stamped with the hash of the kernel sources the capture was taken from --
bench.py nulls the field when the sources have changed since.   usage: make_traffic.py <report.ncu-rep> [...]"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
NAMES = (("loop_fused_kernel", lambda n: "rx_fast_kernel<" in n and ", 1>(" in n),  # the last template argument is TXF
         ("rx_fast_kernel", lambda n: "rx_fast_kernel<64" in n and ", 0>(" in n), ("tx_rect_kernel", lambda n: "tx_rect_fast_kernel" in n),
         ("loop_fused_dec_kernel", lambda n: "rx_dec_kernel<" in n and re.search(r"1>\(", n) is not None),  # rx_dec_kernel<..., TXF = 1>
         ("rx_dec_kernel", lambda n: "rx_dec_kernel<" in n))
out, seen = {}, {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        tot = sum(float(r[hdr.index(k)]) * UNIT[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        for key, match in NAMES:
            if match(name):
                seen.setdefault(key, []).append(tot)
                break
for k, v in seen.items():
    out[k] = int(sum(v) / len(v))
out["source_sha"] = bench.source_sha()
out["_source"] = "ncu --set full --clock-control none (tools/gpu_final.sh), dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches; reports: " + ", ".join(os.path.basename(a) for a in sys.argv[1:])
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
