#!/usr/bin/env bash
# End-to-end pipeline variants: where the D2H copies are issued, number of hardware connections, contexts.
set -u
TAG="${1:-e2e}"
OUT="gpurun_out/e2e_${TAG}"
mkdir -p "$OUT"
run() { local name="$1"; shift; echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"; env E2E_STAGING=8 "$@" > "$OUT/$name.log" 2>&1; grep "ms/step" "$OUT/$name.log" | tee -a "$OUT/session.log"; }
run ranges8_depth3 python tools/e2e_profile4.py 64 8 3
run ranges0_depth3 env JXG_D2H_RANGES=0 python tools/e2e_profile4.py 64 8 3
run ranges8_conn32 env CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/e2e_profile4.py 64 8 3
run ranges0_conn32 env CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_D2H_RANGES=0 python tools/e2e_profile4.py 64 8 3
run ranges0_conn32_depth4 env CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_D2H_RANGES=0 python tools/e2e_profile4.py 64 10 4
run ranges2_conn32 env CUDA_DEVICE_MAX_CONNECTIONS=32 JXG_D2H_RANGES=2 python tools/e2e_profile4.py 64 8 3
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
