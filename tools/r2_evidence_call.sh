#!/usr/bin/env bash
# Round 2, second session: the whole GPU suite, smoke, the bench line, its launch list, and `ncu --set full` of K9 -- ONE gpurun call.
#   gpurun --timeout 2400 -- 'bash tools/r2_evidence_call.sh'
set -u
out=gpurun_out/r2_evidence
mkdir -p "$out"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > "$out/gputests.log" 2>&1
echo "tests rc=$?" | tee -a "$out/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 10 --warmup 3 > "$out/bench_n1.json" 2> "$out/bench_n1.err"
rc=$?
echo "bench rc=$rc" | tee -a "$out/summary.txt"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference_arm.json" 2> "$out/bench_reference_arm.err"
echo "reference arm rc=$?" | tee -a "$out/summary.txt"
if [ "$rc" = 0 ]; then
  # launch list of the same command (2 timed steps): per-launch gpu__time_duration
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$out/launches_r2.csv" \
      python bench.py --steps 2 --warmup 3 --no-cpu > "$out/bench_under_ncu.log" 2>&1
  echo "launch list rc=$?" | tee -a "$out/summary.txt"
fi
# K9: plain, then one full capture of a 40-step launch
timeout 300 python tools/gcn_probe.py > "$out/gcn_plain.log" 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gcn_steps_kernel -s 1 -c 1 -o "$out/gcn" \
    python tools/gcn_probe.py 40 > "$out/gcn_ncu.log" 2>&1
echo "gcn ncu rc=$?" | tee -a "$out/summary.txt"
if [ -f "$out/gcn.ncu-rep" ]; then
  ncu -i "$out/gcn.ncu-rep" --page raw --csv > "$out/gcn_raw.csv" 2>/dev/null
  ncu -i "$out/gcn.ncu-rep" --page source --csv > "$out/gcn_source.csv" 2>/dev/null
  ncu -i "$out/gcn.ncu-rep" --page details > "$out/gcn_details.txt" 2>/dev/null
  rm -f "$out/gcn.ncu-rep"
fi
tail -n 4 "$out/gputests.log" "$out/gcn_plain.log"
ls -la "$out"
