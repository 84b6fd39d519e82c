#!/usr/bin/env bash
# Session S: output copies enqueued by a host thread once their filter range is done (no stream parked on an event wait).
set -u
OUT=gpurun_out/session_r02s
mkdir -p "$OUT"
JXG_D2H_HOST_ORDERED=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests (host-ordered) rc=$?" | tee -a "$OUT/session.log"
tail -2 "$OUT/tests.log" | tee -a "$OUT/session.log"
run() {  # name depth env...
  local name=$1; shift
  local depth=$1; shift
  echo "=== $name depth=$depth ($(date +%T))" | tee -a "$OUT/session.log"
  env "$@" E2E_STAGING=8 timeout 300 python tools/e2e_profile4.py 64 24 $depth > "$OUT/$name.log" 2>&1
  grep -h "ms/step\|main thread" "$OUT/$name.log" | tee -a "$OUT/session.log"
}
run host_d5_marks 5 JXG_D2H_HOST_ORDERED=1 E2E_MARKS=1
run host_d5 5 JXG_D2H_HOST_ORDERED=1
run dev_d5 5
run host_d4 4 JXG_D2H_HOST_ORDERED=1
run host_d6 6 JXG_D2H_HOST_ORDERED=1
run host_d5_r4 5 JXG_D2H_HOST_ORDERED=1 JXG_D2H_RANGES=4
run dev_d5_b 5
run host_d5_b 5 JXG_D2H_HOST_ORDERED=1
timeout 600 python bench.py --config 3 --steps 4 --warmup 3 > "$OUT/bench_config3.log" 2>&1; grep -h '^{' "$OUT/bench_config3.log" | cut -c1-400 | tee -a "$OUT/session.log"
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
