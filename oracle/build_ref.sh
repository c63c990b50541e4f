#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's own hot-path sources from where they lie under
# /root/reference (never copied) into oracle/_ref/libexahype_ref.so.  The reference has no build
# system for these files; this is the g++ line of Unit test/correctness_test.sbatch:24 reduced to
# the two translation units that do not need Peano.  -O0/-O2 give identical bits; contraction is off
# to match a plain x86-64 g++ build.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${EXAHYPE_REFERENCE_DIR:-/root/reference}/Unit test"
if [ ! -f "$ref/test.cpp" ]; then
  echo "reference sources not present at $ref; keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$here/_ref"
"${OCXX:-$( [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++ )}" -std=c++17 -O2 -ffp-contract=off -fPIC -shared -Wl,-Bsymbolic \
    -I"$ref" "$here/ref_shim.cpp" "$ref/test.cpp" "$ref/Functions.cpp" \
    -o "$here/_ref/libexahype_ref.so"
# the same sources as a caller who wants speed would build them (-O3 -march=native, contraction left to the compiler):
# only ever TIMED (bench.py cpu_baseline.reference_compiled), never compared bit for bit
"${OCXX:-$( [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++ )}" -std=c++17 -O3 -march=native -fPIC -shared -Wl,-Bsymbolic \
    -I"$ref" "$here/ref_shim.cpp" "$ref/test.cpp" "$ref/Functions.cpp" \
    -o "$here/_ref/libexahype_ref_fast.so"
echo "built $here/_ref/libexahype_ref.so and libexahype_ref_fast.so"
