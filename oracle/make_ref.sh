#!/bin/bash
# oracle/make_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Builds oracle/_ref/: the UNMODIFIED reference modules of the hot path, taken from where they
# lie under /root/reference, so that the reference itself -- not the port in gnk_oracle.py -- can be timed on the GPU
# box's host cores (bench.py --impl reference, cpu_baseline.kind = "reference") and used to pin the oracle.
# The reference is pure Python, so "building" it is copying the files that make up the path; nothing is edited.
# oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so it
# travels to the GPU box like the built .so files.  /root/reference does not exist there; this script is a no-op then.
#
#   bash oracle/make_ref.sh            # run by __graft_entry__.build() when /root/reference is present
set -euo pipefail
REF="${GNK_REFERENCE_DIR:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "make_ref: $REF not present; keeping $OUT as it is" >&2
  exit 0
fi
mkdir -p "$OUT"
# the modules on the path: solver, Krylov state, line search, second solver (CGLS), result record, the two problems,
# and the harness that the experiment scripts call the solvers through
# ... plus the Bratu experiment script (workload driver of config 1), which tests/ runs UNCHANGED on top
# of the B200 package (install_flat_names(), matplotlib mocked) to show that the reference's own driver drops in
FILES="gauss_newton_krylow.py krylow.py armijo_goldstein.py gauss_newton.py regression_result.py \
       bratu_pde_problem.py rosenbrock_problem.py benchmark.py bratu_pde_test.py"
for f in $FILES; do
  install -m 0644 "$REF/$f" "$OUT/$f"
done
( cd "$REF" && sha256sum $FILES ) > "$OUT/SHA256SUMS"
echo "make_ref: $(ls "$OUT"/*.py | wc -l) reference modules -> $OUT"
