#!/usr/bin/env bash
# Session N: copy-engine probe; end-to-end with the blob uploaded by a kernel instead of cudaMemcpyAsync.
set -u
OUT=gpurun_out/session_r02n
mkdir -p "$OUT"
timeout 200 python tools/copy_engine_probe.py 2>&1 | tee -a "$OUT/session.log"
JXG_UPLOAD_KERNEL=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests (upload kernel) rc=$?" | tee -a "$OUT/session.log"
tail -2 "$OUT/tests.log" | tee -a "$OUT/session.log"
run() {  # name depth env...
  local name=$1; shift
  local depth=$1; shift
  echo "=== $name depth=$depth ($(date +%T))" | tee -a "$OUT/session.log"
  env "$@" E2E_STAGING=8 E2E_MARKS=1 timeout 300 python tools/e2e_profile4.py 64 16 $depth > "$OUT/$name.log" 2>&1
  grep -h "ms/step" "$OUT/$name.log" | tee -a "$OUT/session.log"
}
run upk_d3 3 JXG_UPLOAD_KERNEL=1
run upk_d4 4 JXG_UPLOAD_KERNEL=1
run upk_d5 5 JXG_UPLOAD_KERNEL=1
run upk_d4_r4 4 JXG_UPLOAD_KERNEL=1 JXG_D2H_RANGES=4
run upk_d5_own 5 JXG_UPLOAD_KERNEL=1 JXG_D2H_SHARED=0
run memcpy_d5 5
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
