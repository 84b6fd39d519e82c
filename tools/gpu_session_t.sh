#!/usr/bin/env bash
# Session T (short): full GPU suite + bench lines with the final defaults. Every step has its own timeout.
set -u
OUT=gpurun_out/session_r02t
mkdir -p "$OUT"
step() { local name="$1" limit="$2"; shift 2; echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"; timeout -k 5 "$limit" "$@" > "$OUT/$name.log" 2>&1; echo "    exit $?" | tee -a "$OUT/session.log"; tail -n 2 "$OUT/$name.log" | cut -c1-400 >> "$OUT/session.log"; }
step tests 240 python -m pytest tests -m gpu -q -x
step bench 200 python bench.py --steps 16 --warmup 5
step bench_reference 120 python bench.py --impl reference --steps 3 --warmup 1
step bench_config5 120 python bench.py --config 5 --steps 4 --warmup 3
step bench_config3 200 python bench.py --config 3 --steps 4 --warmup 3
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
