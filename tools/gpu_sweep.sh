#!/usr/bin/env bash
# Sweep of resident batches x entropy lanes per warp on the bench workload (device-resident value only matters here).
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_sweep.sh <tag>'
set -u
TAG="${1:-sweep}"
OUT="gpurun_out/sweep_${TAG}"
mkdir -p "$OUT"
for cfg in ${SWEEP:-"4:4:2 4:4:3 4:4:5 4:3:3 4:3:5 8:8:3 8:8:5"}; do
  IFS=: read S L D <<< "$cfg"
  echo "=== S=$S lanes=$L inflight=$D ($(date +%T))" | tee -a "$OUT/session.log"
  JXG_ENTROPY_S=$S JXG_ENTROPY_LANES=$L JXG_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps $((D*3)) --warmup $D --inflight $D --cpu-sample-frames 1 > "$OUT/bench_S${S}_L${L}_D${D}.log" 2>&1
  grep -h '^{' "$OUT/bench_S${S}_L${L}_D${D}.log" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   value %.0f MP/s, %.2f ms/step, single %.1f ms, stages %s' % (d['value'], d['ms_per_step'], d['config']['single_batch_ms'], {k:round(v,1) for k,v in d['config']['stage_ms_single_batch'].items() if v>0.1}))" | tee -a "$OUT/session.log"
done
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
