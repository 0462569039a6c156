#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the plain-C restatement (oracle/cloudini_oracle.c) into oracle/_build/.
# -ffp-contract=off: `v * mul` must stay a correctly rounded IEEE product (no FMA), like _mm_mul_ps in the reference.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$HERE/_build"
gcc -std=gnu11 -O2 -ffp-contract=off -fPIC -shared -Wall -Wextra -Wno-unused-parameter \
    "$HERE/cloudini_oracle.c" -lm -o "$HERE/_build/libcloudini_oracle.so"
echo "built $HERE/_build/libcloudini_oracle.so"
