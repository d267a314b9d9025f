#!/bin/bash
# Pins the oracle to the REAL reference — on a machine that has cargo (this repository's image does not).
#
#   FRAVE_SRC=/path/to/pagmerek/frave bash oracle/ref_harness/run.sh
#
# Copies the reference tree to a scratch directory (the reference sources are never copied into this
# repository), injects oracle/ref_harness/fri_kat.rs as a test module of libfri, runs it, and writes the
# digest lines to oracle/_ref/kat_digests.jsonl (git-ignored).  `python -m pytest tests/test_oracle.py -k
# reference_digests` then compares them with the C oracle and the plan's emission order.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${FRAVE_SRC:-/root/reference}"
command -v cargo >/dev/null || { echo "cargo not found: the reference cannot be built here (see oracle/ref_harness/README.md)" >&2; exit 3; }
WORK="$(mktemp -d)"
trap 'rm -rf "$WORK"' EXIT
cp -r "$SRC/." "$WORK/"
cp "$HERE/fri_kat.rs" "$WORK/crates/libfri/src/stages/fri_kat.rs"
printf '\n#[cfg(test)]\nmod fri_kat;\n' >> "$WORK/crates/libfri/src/stages/mod.rs"
mkdir -p "$HERE/../_ref"
(cd "$WORK" && cargo test --release -p libfri fri_kat -- --nocapture) | tee "$WORK/out.txt"
grep '^FRI_KAT ' "$WORK/out.txt" | sed 's/^FRI_KAT //' > "$HERE/../_ref/kat_digests.jsonl"
echo "wrote $(wc -l < "$HERE/../_ref/kat_digests.jsonl") digest line(s) to oracle/_ref/kat_digests.jsonl"
