#!/usr/bin/env bash
# Session R: full GPU suite, launch list of one 64-frame step (ncu), bench lines of every config, reference arm,
# end-to-end timeline on the host clock.
set -u
OUT=gpurun_out/session_r02r
mkdir -p "$OUT"
step() { local name="$1" limit="$2"; shift 2; echo "=== $name ($(date +%T))" | tee -a "$OUT/session.log"; timeout "$limit" "$@" > "$OUT/$name.log" 2>&1; echo "    exit $?" | tee -a "$OUT/session.log"; tail -n 3 "$OUT/$name.log" | cut -c1-300 >> "$OUT/session.log"; }
step tests 1200 python -m pytest tests -m gpu -q -x --durations=8
step smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step stages 200 python tools/stage_times.py 64
step launches 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file "$OUT/launches.csv" python tools/stage_times.py 64
step e2e_marks 300 env E2E_STAGING=8 E2E_MARKS=1 python tools/e2e_profile4.py 64 16 5
step bench 600 python bench.py --steps 16 --warmup 5
step bench_reference 600 python bench.py --impl reference --steps 3 --warmup 1
step bench_config5 600 python bench.py --config 5 --steps 4 --warmup 3
step bench_config3 600 python bench.py --config 3 --steps 4 --warmup 3
step bench_config4 900 python bench.py --config 4 --steps 4 --warmup 3
echo "=== done ($(date +%T))" | tee -a "$OUT/session.log"
