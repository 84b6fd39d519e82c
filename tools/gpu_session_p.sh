#!/usr/bin/env bash
# Session P: 16-bit outputs + palette tests, ncu capture of the Modular decode kernel, what the device does after drain().
set -u
OUT=gpurun_out/session_r02p
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_modular.py tests/test_gpu_pipeline.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$OUT/session.log"
tail -3 "$OUT/tests.log" | tee -a "$OUT/session.log"
for TK in 1 0 2; do timeout 300 python tools/modular_once.py 8 4096 $TK 3 2>&1 | tail -1 | tee -a "$OUT/session.log"; done
MODULAR_LANES=4 timeout 300 python tools/modular_once.py 8 4096 1 3 2>&1 | tail -1 | tee -a "$OUT/session.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_modular_decode -c 1 -f -o "$OUT/modular_decode" python tools/modular_once.py 8 4096 1 0 > "$OUT/ncu_modular.log" 2>&1; echo "ncu rc=$?" | tee -a "$OUT/session.log"
E2E_STAGING=8 timeout 300 python tools/e2e_profile4.py 64 16 5 > "$OUT/e2e_d5_nomarks.log" 2>&1; grep -h "ms/step\|main thread\|second sync" "$OUT/e2e_d5_nomarks.log" | tee -a "$OUT/session.log"
E2E_STAGING=8 timeout 300 python tools/e2e_profile4.py 64 16 4 > "$OUT/e2e_d4_nomarks.log" 2>&1; grep -h "ms/step\|main thread\|second sync" "$OUT/e2e_d4_nomarks.log" | tee -a "$OUT/session.log"
timeout 600 python bench.py --steps 10 --warmup 5 --inflight 5 > "$OUT/bench_inflight5.log" 2>&1; grep -h '^{' "$OUT/bench_inflight5.log" | tee -a "$OUT/session.log"
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
