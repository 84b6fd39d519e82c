#!/usr/bin/env bash
# Shared-memory carve-out x stream structure on the bench workload (device-resident leg only).
set -u
TAG="${1:-carve}"
OUT="gpurun_out/carve_${TAG}"
mkdir -p "$OUT"
for cfg in "-1 0 3" "100 0 3" "100 0 2" "100 1 2" "100 1 3" "75 0 3" "75 1 3" "100 0 4"; do
  set -- $cfg
  C=$1; ST=$2; D=$3
  echo "=== carveout=$C stage_streams=$ST inflight=$D ($(date +%T))" | tee -a "$OUT/session.log"
  JXG_CARVEOUT=$C JXG_STAGE_STREAMS=$ST JXG_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps $((D*3)) --warmup $D --inflight $D --cpu-sample-frames 1 > "$OUT/bench_C${C}_T${ST}_D${D}.log" 2>&1
  grep -h '^{' "$OUT/bench_C${C}_T${ST}_D${D}.log" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   value %.0f MP/s, %.2f ms/step, single %.1f ms, stages %s' % (d['value'], d['ms_per_step'], d['config']['single_batch_ms'], {k:round(v,1) for k,v in d['config']['stage_ms_single_batch'].items() if v>0.1}))" | tee -a "$OUT/session.log"
done
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
