#!/usr/bin/env bash
# usage: tools/lib_variants.sh name1 name2 ...  — quick bench of cloudini_b200/lib/var/<name>.so builds (development helper)
bash tools/quick_bench.sh base
for n in "$@"; do
  CLDN_B200_LIB=$PWD/cloudini_b200/lib/var/$n.so bash tools/quick_bench.sh $n
done
