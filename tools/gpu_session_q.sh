#!/usr/bin/env bash
# Session Q: Modular channel loop v2 (tests + timings), bench with the new defaults.
set -u
OUT=gpurun_out/session_r02q
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_modular.py -m gpu -x -q > "$OUT/tests.log" 2>&1; echo "modular tests rc=$?" | tee -a "$OUT/session.log"
tail -3 "$OUT/tests.log" | tee -a "$OUT/session.log"
for TK in 1 0 2 3; do timeout 300 python tools/modular_once.py 8 4096 $TK 3 2>&1 | tail -1 | tee -a "$OUT/session.log"; done
for L in 2 4; do MODULAR_LANES=$L timeout 300 python tools/modular_once.py 8 4096 1 3 2>&1 | tail -1 | sed "s/^/lanes=$L: /" | tee -a "$OUT/session.log"; done
timeout 600 python bench.py --steps 16 --warmup 5 > "$OUT/bench.log" 2>&1; grep -h '^{' "$OUT/bench.log" | tee -a "$OUT/session.log"
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 > "$OUT/bench5.log" 2>&1; grep -h '^{' "$OUT/bench5.log" | tee -a "$OUT/session.log"
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
