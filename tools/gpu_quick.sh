#!/usr/bin/env bash
# Short GPU session: a subset of the GPU tests, one bench line, the end-to-end timeline and sweeps of the entropy
# schedule knobs. Usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
set -u
TAG="${1:-quick}"
OUT="gpurun_out/quick_${TAG}"
mkdir -p "$OUT"
step() { local name="$1" limit="$2"; shift 2; echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"; timeout "$limit" "$@" > "$OUT/$name.log" 2>&1; echo "    exit $?" | tee -a "$OUT/session.log"; tail -n 3 "$OUT/$name.log" >> "$OUT/session.log"; }
step tests 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py tests/test_gpu_pipeline.py -m gpu -q -x -k "not config4"
step stages_default 200 python tools/stage_times.py 64
for solo in 0.5 0.75 1.1; do
  step "stages_solo_${solo}" 200 env JXG_ENTROPY_SOLO=$solo python tools/stage_times.py 64
done
step stages_solo_off_perlane1 200 env JXG_ENTROPY_SOLO=1.1 JXG_ENTROPY_DUO=1.1 JXG_ENTROPY_PER_LANE=1.0 python tools/stage_times.py 64
step stages_perlane_2.5 200 env JXG_ENTROPY_PER_LANE=2.5 python tools/stage_times.py 64
step stages_S8 200 env JXG_ENTROPY_S=8 python tools/stage_times.py 64
step bench 400 python bench.py --steps 8 --warmup 4
step e2e 300 python tools/e2e_profile4.py 64 8 3
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
