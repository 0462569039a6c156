#!/usr/bin/env bash
# Development aid: build tests/cusim (the CPU emulation of the kernels; test infrastructure only) and run a selection of
# the `-m gpu` tests against it.   tools/cusim_run.sh tests/test_gpu_fast_paths.py -k mixed
set -eu
cd "$(dirname "$0")/.."
LIB=$(python -c "import sys; sys.path.insert(0,'tests/cusim'); import build_cusim; print(build_cusim.build())" | tail -1)
[ -f "$LIB" ] || { echo "cusim build failed"; exit 1; }
CLDN_B200_LIB=$LIB CLDN_B200_ALLOW_EMULATION=tests-only python -m pytest -x -q -m gpu -p no:cacheprovider "$@"
