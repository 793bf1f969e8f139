#!/usr/bin/env bash
# Round 2, first GPU call: what round 1 left unverified on hardware, in ONE gpurun call.
#   gpurun --timeout 900 -- 'bash tools/r2_first_gpu_call.sh'
# Everything lands in gpurun_out/r2_first/.  Each step runs under its own timeout; ncu only after the plain run exited 0.
set -u
out=gpurun_out/r2_first
mkdir -p "$out"
# 1. K8 (CUNE) GPU tests: XPASS = the kernel matches the reference loop on hardware -> drop the xfail marks
timeout 300 python -m pytest tests/test_zz_cune.py -m gpu -q -rxX > "$out/cune_tests.log" 2>&1
echo "cune tests rc=$?" | tee -a "$out/summary.txt"
# 2. smoke (prints the K8 line after the asserted legs)
timeout 300 python __graft_entry__.py smoke > "$out/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$out/summary.txt"
grep -h "K8 first hardware run" "$out/smoke.log" | tee -a "$out/summary.txt"
# 3. K8 at scale: one Hogwild epoch at config C2's shape, timed on the device
timeout 600 python tools/cune_probe.py > "$out/cune_probe.log" 2>&1
rc=$?
echo "cune probe rc=$rc" | tee -a "$out/summary.txt"
tail -n 5 "$out/cune_probe.log" | tee -a "$out/summary.txt"
# 4. ncu of the K8 kernel (only after the probe ran clean)
if [ "$rc" = 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cune_sgd_kernel -c 1 \
      -o "$out/cune_k8" python tools/cune_probe.py --small > "$out/ncu.log" 2>&1
  echo "ncu rc=$?" | tee -a "$out/summary.txt"
fi
