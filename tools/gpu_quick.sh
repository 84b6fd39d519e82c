#!/usr/bin/env bash
# Short GPU session: GPU tests, stage times, one bench line, the end-to-end timeline with the host-side phase trace of
# jxg_batch_run. Usage: gpurun --timeout 1200 -- 'bash tools/gpu_quick.sh <tag>'   (TESTS="..." selects test files)
set -u
TAG="${1:-quick}"
OUT="gpurun_out/quick_${TAG}"
mkdir -p "$OUT"
step() { local name="$1" limit="$2"; shift 2; echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"; timeout "$limit" "$@" > "$OUT/$name.log" 2>&1; echo "    exit $?" | tee -a "$OUT/session.log"; tail -n 4 "$OUT/$name.log" >> "$OUT/session.log"; }
step tests 900 python -m pytest ${TESTS:-tests} -m gpu -q -x --durations=5
step stages 200 python tools/stage_times.py 64
step bench 400 python bench.py --steps 8 --warmup 4
step e2e 300 env JXG_TRACE_RUN=1 python tools/e2e_profile4.py 64 6 3
if [[ "${LAUNCHES:-0}" == 1 ]]; then
  step launches 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file "$OUT/launches.csv" python tools/stage_times.py 64
fi
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
