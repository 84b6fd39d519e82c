#!/usr/bin/env bash
# Session M: first-in-first-out D2H stream shared by all contexts of the device against one copy stream per context.
set -u
OUT=gpurun_out/session_r02m
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/session.log"
tail -2 "$OUT/tests.log" | tee -a "$OUT/session.log"
run() {  # name depth env...
  local name=$1; shift
  local depth=$1; shift
  echo "=== $name depth=$depth ($(date +%T))" | tee -a "$OUT/session.log"
  env "$@" E2E_STAGING=8 E2E_MARKS=1 timeout 300 python tools/e2e_profile4.py 64 16 $depth > "$OUT/$name.log" 2>&1
  grep -h "ms/step" "$OUT/$name.log" | tee -a "$OUT/session.log"
}
run fifo_d3_s4 3 JXG_ENTROPY_S=4
run fifo_d4_s4 4 JXG_ENTROPY_S=4
run fifo_d4_s8 4 JXG_ENTROPY_S=8
run fifo_d5_s8 5 JXG_ENTROPY_S=8
run fifo_d6_s8 6 JXG_ENTROPY_S=8
run fifo_d4_s8_r4 4 JXG_ENTROPY_S=8 JXG_D2H_RANGES=4
run fifo_d4_s8_r1 4 JXG_ENTROPY_S=8 JXG_D2H_RANGES=1
run own_d4_s8 4 JXG_ENTROPY_S=8 JXG_D2H_SHARED=0
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
