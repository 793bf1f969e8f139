#!/usr/bin/env bash
# ncu --set full of rank_tc_kernel on one wave (18 944 users x 2 M tracks, d = 64); raw / source pages as CSV.
set -u
out=gpurun_out/r2_rank
mkdir -p "$out"
timeout 300 python tools/rank_probe.py 18944 2000000 3 > "$out/rank_plain.log" 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 1 -c 1 -o "$out/rank_tc" \
    python tools/rank_probe.py 18944 2000000 1 > "$out/rank_ncu.log" 2>&1
echo "rank ncu rc=$?" | tee -a "$out/summary.txt"
if [ -f "$out/rank_tc.ncu-rep" ]; then
  ncu -i "$out/rank_tc.ncu-rep" --page raw --csv > "$out/rank_tc_raw.csv" 2>/dev/null
  ncu -i "$out/rank_tc.ncu-rep" --page source --csv > "$out/rank_tc_source.csv" 2>/dev/null
  rm -f "$out/rank_tc.ncu-rep"
fi
tail -n 4 "$out/rank_plain.log"
