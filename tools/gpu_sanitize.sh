#!/bin/bash
# compute-sanitizer over the small-shape GPU parity tests (every kernel family: fused loopback with TMEM, fast RX 64/129 taps,
# shaped/rect TX, generic kernels, AWGN, bank, phasor scan, PLL lock).  Full-size tests are skipped (the tools slow kernels 10-100x).
# Output: gpurun_out/sanitize_<tool>.log (summaries are copied to profiles/ by hand).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
SKIP='not full_size and not across_2p24 and not torch and not large_sample_index'
for tool in memcheck racecheck synccheck initcheck; do
  sel="$SKIP"
  # racecheck / synccheck / initcheck: the kernels with shared-memory phases and barriers are what matters; keep the run short
  if [ "$tool" != memcheck ]; then sel="$SKIP and (fused or rx_fast or loopback_all or shaped_fast or stateful_tx or lock_phase or noisy or bank or demodulate_bin)"; fi
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no --padding 32"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 20 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest tests -m gpu -q -x -k "$sel" -p no:cacheprovider > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_rc.txt
  tail -3 gpurun_out/sanitize_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$tool.log | tail -3
done
