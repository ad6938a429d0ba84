#!/bin/bash
# copies the evidence of tools/gpu_final.sh (gpurun_out/fin_*) into profiles/ under the round's names and writes the ncu summaries
# usage: tools/collect_profiles.sh r02   (run in the repo root, after the gpurun call came back)
R=${1:-r02}; O=gpurun_out/fin; P=profiles
cp ${O}_bench.json $P/${R}_final_bench_1gpu.json
cp ${O}_bench_ref.json $P/${R}_final_bench_reference_arm.json
cp ${O}_configs.jsonl $P/${R}_final_configs_1gpu.jsonl
cp ${O}_frontend.jsonl $P/${R}_final_frontend_rows.jsonl
cp ${O}_ber_sweep_1gpu.json $P/${R}_ber_sweep_1gpu.json
cp ${O}_pcie_floor_1gpu.txt $P/${R}_pcie_floor_1gpu.txt
cp ${O}_packed_sweep.txt $P/${R}_packed_sweep.txt
cp ${O}_check_sqrt.txt $P/${R}_check_sqrt.txt
[ -f ${O}_check_sincos_dev.txt ] && cp ${O}_check_sincos_dev.txt $P/${R}_check_sincos_dev.txt
cp ${O}_launches.csv $P/${R}_final_launches.csv
python - <<PY > $P/${R}_final_launches.txt
import csv, collections
rows = [r for r in csv.reader(open("${O}_launches.csv", errors="replace")) if len(r) > 10]
hdr = rows[0]; k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: t = float(r[v].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(r[k], [0, 0.0]); a[0] += 1; a[1] += t
print("# launch list of \`python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs c1\` (ncu --metrics gpu__time_duration.sum --clock-control none, first 200 launches): launches and mean us per kernel")
for n, (c, t) in agg.items():
    print(f"{c:4d} x {t / c / 1e3:9.1f} us  {n[:100]}")
PY
{
echo "# ncu --set full --clock-control none --import-source on, one B200 (gpurun), tools/gpu_final.sh: captured after the same command"
echo "# (python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs c1, workload C2) exited 0 without ncu."
echo "# rx_fast_kernel<..., 128, 4, 4, 3, 64, 1> = the FUSED LOOPBACK kernel (one launch per bench step); tx_rect_fast_kernel and rx_fast_kernel<..., 64, 8, 4, 5, 64, 0> = the two-kernel path timed beside it."
echo
python tools/ncu_summary.py ${O}_c2_prof.ncu-rep
} > $P/${R}_final_c2_ncu.txt
{
echo "# ncu --set full, same conditions, python tools/bench_configs.py c1l: rx_dec_kernel<64, 0, 1, 1> = the FUSED LOOPBACK at the reference's default rates (sps 45), 4096 frames x 65520 samples, one kernel"
echo
python tools/ncu_summary.py ${O}_c1l_prof.ncu-rep
echo
echo "# ncu --set full, same conditions, python tools/bench_configs.py c1: rx_dec_kernel<64, 0, 1, 0> (the unfused RX) at the same shape"
echo
python tools/ncu_summary.py ${O}_c1_prof.ncu-rep
echo
echo "# ncu --set full, same conditions, python tools/bench_configs.py c2n: the noisy fast RX kernel (C2 + AWGN 6 dB added while loading)"
echo
python tools/ncu_summary.py ${O}_c2n_prof.ncu-rep
} > $P/${R}_final_c1_c2n_ncu.txt
python tools/sass_histogram.py > $P/${R}_sass_histogram.txt 2>/dev/null || true
