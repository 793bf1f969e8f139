#!/usr/bin/env bash
# Round 2: the bench line of one GPU, then `ncu --set full` of the three SGD-family kernels, in ONE gpurun call
# (all ncu runs of a call count as one; each only after the same command exited 0 without ncu).
#   gpurun --timeout 1500 -- 'bash tools/r2_ncu_call.sh'
# Everything lands in gpurun_out/r2_ncu/.
set -u
out=gpurun_out/r2_ncu
mkdir -p "$out"
timeout 600 python bench.py --steps 10 --warmup 3 > "$out/bench_n1.json" 2> "$out/bench_n1.err"
echo "bench rc=$?" | tee -a "$out/summary.txt"
# config C2, d = 64: bpr_sgd_blk_kernel<2, false, false>
export PROBE_EPOCHS=3
timeout 300 python tools/sgd_probe.py > "$out/c2_plain.log" 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bpr_sgd_blk_kernel -s 10 -c 1 -o "$out/sgd_c2" \
    python tools/sgd_probe.py > "$out/c2_ncu.log" 2>&1
echo "c2 ncu rc=$?" | tee -a "$out/summary.txt"
# config C3's shard (1 of 8), d = 128: bpr_sgd_blk_kernel<4, false, false>; Q = 1 GB > L2
export PROBE_SIZE=1250000,2000000,125000000,128
timeout 300 python tools/sgd_probe.py > "$out/c3_plain.log" 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bpr_sgd_blk_kernel -s 10 -c 1 -o "$out/sgd_c3" \
    python tools/sgd_probe.py > "$out/c3_ncu.log" 2>&1
echo "c3 ncu rc=$?" | tee -a "$out/summary.txt"
unset PROBE_SIZE
# K8 (CUNE), 100 K users x 20 K tracks x 5 M events, d = 64
timeout 300 python tools/cune_probe.py > "$out/cune_plain.log" 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cune_sgd_kernel -s 1 -c 1 -o "$out/cune" \
    python tools/cune_probe.py > "$out/cune_ncu.log" 2>&1
echo "cune ncu rc=$?" | tee -a "$out/summary.txt"
# the reports are ~25 MB each and gpurun brings back 64 MiB: keep their raw and source pages as CSV, drop the reports
for r in sgd_c2 sgd_c3 cune; do
  if [ -f "$out/$r.ncu-rep" ]; then
    ncu -i "$out/$r.ncu-rep" --page raw --csv > "$out/${r}_raw.csv" 2>/dev/null
    ncu -i "$out/$r.ncu-rep" --page source --csv > "$out/${r}_source.csv" 2>/dev/null
    ncu -i "$out/$r.ncu-rep" --page details > "$out/${r}_details.txt" 2>/dev/null
    rm -f "$out/$r.ncu-rep"
  fi
done
tail -n 3 "$out/c2_plain.log" "$out/c3_plain.log" "$out/cune_plain.log"
ls -la "$out"
