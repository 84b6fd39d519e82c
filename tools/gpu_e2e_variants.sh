#!/usr/bin/env bash
# End-to-end pipeline variants with the device timeline of every batch (E2E_MARKS): number of contexts, D2H ranges.
set -u
TAG="${1:-e2e}"
OUT="gpurun_out/e2e_${TAG}"
mkdir -p "$OUT"
run() { local name="$1"; shift; echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"; env E2E_STAGING=8 E2E_MARKS=1 "$@" > "$OUT/$name.log" 2>&1; grep "ms/step" "$OUT/$name.log" | tee -a "$OUT/session.log"; }
run ranges8_depth3 python tools/e2e_profile4.py 64 8 3
run ranges8_depth2 python tools/e2e_profile4.py 64 8 2
run ranges2_depth3 env JXG_D2H_RANGES=2 python tools/e2e_profile4.py 64 8 3
run ranges0_depth2 env JXG_D2H_RANGES=0 python tools/e2e_profile4.py 64 8 2
echo "=== bench ($(date +%T))" | tee -a "$OUT/session.log"
timeout 400 python bench.py --steps 9 --warmup 3 > "$OUT/bench.log" 2>&1; grep -h '^{' "$OUT/bench.log" | cut -c1-200 | tee -a "$OUT/session.log"
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
