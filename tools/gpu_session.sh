#!/usr/bin/env bash
# One gpurun call's worth of measurements, so that a round spends its GPU minutes on one box acquisition:
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r02a'
#
# Everything lands in gpurun_out/session_<tag>/ (merged back into the build container). Every step runs under its
# own `timeout` and a failing step does not stop the later ones; the exit code is that of the GPU test suite.
# Steps (skip any with SKIP="tests ncu_full ..."):
#   tests      python -m pytest tests -m gpu -x -q
#   bench      bench.py (our arm) and bench.py --impl reference, N=1, default workload (BASELINE config 2)
#   stages     per-stage device times of one resident batch (tools/stage_times.py)
#   e2e        host-side timeline of PipelinedDecoder (tools/e2e_profile4.py)
#   hostprobe  CPU quota and front-end parse throughput of the box (tools/host_probe.py)
#   sweeps     resident batches 1/2/3, batch of 128 frames, 1080p frames (config 3 shape), libjxl-like LF coding,
#              one 16384^2 EPF-3 frame (config 4)
#   modular    config 5 (tools/bench_modular.py)
#   launches   ncu launch list of a short bench run (per-launch times; shares, not absolutes)
#   ncu_full   ncu --set full of the entropy, transform and filter kernels on a small batch (+ raw CSV pages)
set -u
TAG="${1:-session}"
OUT="gpurun_out/session_${TAG}"
mkdir -p "$OUT"
SKIP=" ${SKIP:-} "
want() { [[ "$SKIP" != *" $1 "* ]]; }
step() {  # step <name> <timeout-seconds> <command...>
  local name="$1" limit="$2"
  shift 2
  echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"
  timeout "$limit" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "    exit $rc" | tee -a "$OUT/session.log"
  tail -n 3 "$OUT/$name.log" >> "$OUT/session.log"
  return $rc
}

nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > "$OUT/gpu.csv" 2>&1
ls -la jxl_rs_b200/libjxgpu.so oracle/liboracle.so synth/libjxlsynth.so > "$OUT/build.log" 2>&1  # prebuilt in the build container; they travel with the snapshot

TESTS_RC=0
if want tests; then
  step tests 1500 python -m pytest tests -m gpu -q --durations=15; TESTS_RC=$?
  step smoke 300 python __graft_entry__.py smoke
fi
if want bench; then
  step bench_reference 400 python bench.py --impl reference --steps 2 --warmup 1
  step bench 600 python bench.py --steps 8 --warmup 4
  grep -h '^{' "$OUT/bench.log" > "$OUT/bench.json" || true
  grep -h '^{' "$OUT/bench_reference.log" > "$OUT/bench_reference.json" || true
fi
if want stages; then step stages 300 python tools/stage_times.py 64; fi
if want e2e; then step e2e 400 python tools/e2e_profile4.py 64 8; fi
if want hostprobe; then step hostprobe 300 python tools/host_probe.py; fi
if want sweeps; then
  step sweep_inflight3 300 python bench.py --steps 6 --warmup 3 --inflight 3
  step sweep_frames128 500 python bench.py --steps 4 --warmup 3 --frames 128
  step config3 500 python bench.py --config 3 --steps 4 --warmup 3
  step config4 600 python bench.py --config 4 --steps 3 --warmup 3
fi
if want modular; then
  step config5 500 python bench.py --config 5 --steps 3 --warmup 3
fi
if want launches; then
  step launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --frames 64
fi
if want ncu_full; then
  # small batch: ncu replays every profiled launch ~40 times
  for k in k_entropy_lean k_idct_small k_dequant_idct k_filters_store; do
    step "ncu_$k" 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 2 -o "$OUT/prof_$k" \
      python tools/stage_times.py 16
    if [[ -f "$OUT/prof_$k.ncu-rep" ]]; then
      ncu -i "$OUT/prof_$k.ncu-rep" --page raw --csv > "$OUT/prof_${k}_raw.csv" 2>/dev/null || true
      ncu -i "$OUT/prof_$k.ncu-rep" --page source --csv > "$OUT/prof_${k}_source.csv" 2>/dev/null || true
      # gpurun merges at most 64 MiB back: keep the CSV pages, drop the report unless asked to keep it
      [[ "${KEEP_REP:-0}" == 1 ]] || rm -f "$OUT/prof_$k.ncu-rep"
    fi
  done
fi
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
exit $TESTS_RC
