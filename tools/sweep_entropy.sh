#!/bin/bash
# Sweeps the persistent-lane knobs of k_entropy_lean; prints the entropy stage time of one batch.
for S in 2 4 8; do for L in 2000 3000 4500 6000 8640; do
  JXG_ENTROPY_S=$S JXG_ENTROPY_LANES=$L python bench.py --steps 2 --warmup 1 --cpu-sample-frames 0 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('S=$S lanes=$L', {k:round(v,2) for k,v in d['config']['stage_ms_single_batch'].items() if v>0.1}, round(d['ms_per_step'],2))"
done; done
