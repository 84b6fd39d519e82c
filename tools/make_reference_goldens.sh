#!/usr/bin/env bash
# Pins decoded pixels against the ACTUAL jxl-rs binary (the step this repository's build image cannot do: it has no
# Rust toolchain). Run on any machine with cargo and a checkout of libjxl/jxl-rs:
#
#   tools/make_reference_goldens.sh /path/to/jxl-rs [extra.jxl ...]
#
# It builds jxl_cli (release), decodes every VarDCT fixture under tests/golden/jxl/ plus three synthetic frames of
# the bench workload to float32 .npy files (jxl_cli/src/enc/numpy.rs: shape frames x H x W x C, sRGB-encoded
# samples in the image's output colour profile, no dither, no rounding) under tests/golden/pixels/, and prints the
# reference's own --speedtest line (jxl_cli/src/main.rs:198-238) for each input. Commit tests/golden/pixels/*.npy
# (or keep them local): tests/test_reference_pixels.py then compares the CPU oracle — and, on a GPU box, the CUDA
# path — against them and the "parity unpinned" note in DESIGN.md can go.
set -euo pipefail
REF="${1:?usage: $0 /path/to/jxl-rs [extra .jxl files]}"
shift || true
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/tests/golden/pixels"
mkdir -p "$OUT"
(cd "$REF" && cargo build --release -p jxl_cli)
CLI="$REF/target/release/jxl_cli"
# synthetic frames of the bench workload (seeds 2000..2002, 1280x720 so that the goldens stay small)
python - "$ROOT" "$OUT" <<'PY'
import os, sys
sys.path.insert(0, sys.argv[1])
import synth
for seed in (2000, 2001, 2002):
    p = os.path.join(sys.argv[2], f"synthetic_{seed}.jxl")
    open(p, "wb").write(synth.encode_synthetic(1280, 720, seed, 0.5, 2, 1, 1, lf_tree=seed & 1))
PY
for f in "$ROOT"/tests/golden/jxl/*.jxl "$OUT"/synthetic_*.jxl "$@"; do
  name="$(basename "$f" .jxl)"
  if "$CLI" "$f" "$OUT/$name.npy" --data-type f32 2> "$OUT/$name.err"; then
    "$CLI" "$f" --speedtest --num-reps 5 --data-type u8 | tail -n 1 | sed "s|^|$name: |"
    rm -f "$OUT/$name.err"
  else
    echo "$name: jxl_cli failed (see $OUT/$name.err)"
  fi
done
