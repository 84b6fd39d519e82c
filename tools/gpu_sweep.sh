#!/usr/bin/env bash
# Sweep of resident batches x entropy lanes per warp on the bench workload (device-resident value only matters here).
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_sweep.sh <tag>'
set -u
TAG="${1:-sweep}"
OUT="gpurun_out/sweep_${TAG}"
mkdir -p "$OUT"
SWEEP="${SWEEP:-4:4:2 4:4:3 4:4:4 4:3:3 8:8:2 8:8:3 4:4:2:0 4:4:3:0}"
for cfg in $SWEEP; do
  IFS=: read S L D ST <<< "$cfg"
  ST="${ST:-1}"  # 4th field 0: one stream per batch instead of the device's two stage streams
  echo "=== S=$S lanes=$L inflight=$D stage_streams=$ST ($(date +%T))" | tee -a "$OUT/session.log"
  JXG_STAGE_STREAMS=$ST JXG_ENTROPY_S=$S JXG_ENTROPY_LANES=$L JXG_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps $((D*3)) --warmup $D --inflight $D --cpu-sample-frames 1 > "$OUT/bench_S${S}_L${L}_D${D}_T${ST}.log" 2>&1
  grep -h '^{' "$OUT/bench_S${S}_L${L}_D${D}_T${ST}.log" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   value %.0f MP/s, %.2f ms/step, single %.1f ms, stages %s' % (d['value'], d['ms_per_step'], d['config']['single_batch_ms'], {k:round(v,1) for k,v in d['config']['stage_ms_single_batch'].items() if v>0.1}))" | tee -a "$OUT/session.log"
done
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
