#!/usr/bin/env bash
# Time the UNMODIFIED reference (probcomp/PClean, Julia) on its three shipped experiments, and dump
# reference-held fixtures that pin the oracle (tests/golden/reference_fixtures/*.json).
#
# Neither the build image nor the GPU box has a `julia` binary or a Julia depot, so
# this script cannot run there; it records that fact and exits 0.  On any machine with Julia >= 1.5
# and network access for `Pkg.instantiate()`:
#
#     PCLEAN_REF=/path/to/PClean baseline/run_reference.sh
#
# writes baseline/reference_times.json ({experiment: {init_s, sweeps_s, total_s, f1...}}) and the
# fixtures; `python -m pytest tests/test_reference_fixtures.py` then compares the oracle with them
# (that test skips while the fixture directory is absent: parity stays "unpinned" until someone runs this).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${PCLEAN_REF:-/root/reference}"
out="$here/reference_times.json"
if ! command -v julia >/dev/null 2>&1; then
  echo '{"status": "reference CPU path: not runnable in this image (no julia binary)"}' > "$out"
  echo "run_reference.sh: julia not found; wrote $out" >&2
  exit 0
fi
if [ ! -d "$ref/src" ]; then
  echo "run_reference.sh: PCLEAN_REF=$ref has no src/ (point it at a checkout of probcomp/PClean)" >&2
  exit 1
fi
fixtures="$here/../tests/golden/reference_fixtures"
mkdir -p "$fixtures"
julia --project="$ref" -e 'using Pkg; Pkg.instantiate()'
julia --project="$ref" "$here/time_experiments.jl" "$ref" "$out"
julia --project="$ref" "$here/dump_fixtures.jl" "$ref" "$fixtures"
echo "run_reference.sh: wrote $out and $fixtures" >&2
