#!/usr/bin/env bash
# Session L: is the end-to-end pipeline held back by stream -> hardware-queue aliasing (CUDA_DEVICE_MAX_CONNECTIONS)?
set -u
OUT=gpurun_out/session_r02l
mkdir -p "$OUT"
run() {  # name, env..., depth
  local name=$1; shift
  local depth=$1; shift
  echo "=== $name depth=$depth ($(date +%T))" | tee -a "$OUT/session.log"
  env "$@" E2E_STAGING=8 E2E_MARKS=1 timeout 300 python tools/e2e_profile4.py 64 16 $depth > "$OUT/$name.log" 2>&1
  grep -h "ms/step" "$OUT/$name.log" | tee -a "$OUT/session.log"
}
run conn32_d4 4 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_d5 5 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_d4_s8 4 CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_ENTROPY_S=8
run conn32_d6_s8 6 CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_ENTROPY_S=8
run conn8_d4_ranges0 4 JXG_D2H_RANGES=0
run conn32_d4_ranges0 4 CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_D2H_RANGES=0
run conn32_d6_ranges0 6 CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_D2H_RANGES=0
run conn8_d4_trace 4 JXG_TRACE_RUN=1
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
# filter CTA size x persistent prefetching kernel (single resident batch, stage times)
for cfg in "512 0" "384 0" "256 0" "512 1" "384 1" "256 1"; do
  set -- $cfg
  echo "=== filters threads=$1 persistent=$2" | tee -a "$OUT/session.log"
  JXG_FILTER_THREADS=$1 JXG_FILTERS_PERSISTENT=$2 timeout 200 python tools/stage_times.py 64 2>&1 | tail -1 | tee -a "$OUT/session.log"
done
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
