#!/usr/bin/env bash
# Session K: persistent prefetching filter kernel (tests + stage times), wide entropy warps x resident batches,
# end-to-end timeline with the H2D / D2H ends.
set -u
OUT=gpurun_out/session_r02k
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/session.log"
tail -3 "$OUT/tests.log" | tee -a "$OUT/session.log"
SWEEP="4:4:3:0 8:8:3:0 8:8:5:0 16:16:3:0 16:16:5:0 32:32:3:0 32:32:5:0 4:4:5:0" bash tools/gpu_sweep.sh r02k
cp gpurun_out/sweep_r02k/session.log "$OUT/sweep.log"
for D in 4 5; do
  echo "=== e2e depth=$D" | tee -a "$OUT/session.log"
  E2E_STAGING=8 E2E_MARKS=1 timeout 300 python tools/e2e_profile4.py 64 16 $D > "$OUT/e2e_d$D.log" 2>&1
  grep -h "ms/step" "$OUT/e2e_d$D.log" | tee -a "$OUT/session.log"
done
