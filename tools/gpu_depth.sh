#!/usr/bin/env bash
# End-to-end leg: contexts in flight x sub-batches per step (host dispatcher + device timelines).
set -u
TAG="${1:-depth}"
OUT="gpurun_out/depth_${TAG}"
mkdir -p "$OUT"
for cfg in "3 1" "4 1" "5 1" "6 1" "6 2" "8 2" "12 4"; do
  set -- $cfg
  D=$1; SP=$2
  echo "=== depth=$D split=$SP ($(date +%T))" | tee -a "$OUT/session.log"
  E2E_SPLIT=$SP E2E_STAGING=8 E2E_MARKS=1 timeout 300 python tools/e2e_profile4.py 64 16 $D > "$OUT/d${D}_s${SP}.log" 2>&1
  grep -h "ms/step" "$OUT/d${D}_s${SP}.log" | tee -a "$OUT/session.log"
done
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
