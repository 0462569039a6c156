#!/usr/bin/env bash
# One `ncu --set full` capture per workload of tools/kernel_tour.py (every shipped kernel exactly once, inside
# cudaProfilerStart/Stop). The .ncu-rep files are ~40 MB each with sources imported and gpurun brings back at most 64 MiB,
# so what travels is extracted on the box: the raw page (all metrics per kernel) and, for the kernels matching SRC_KERNELS,
# the per-source-line page (instructions + stall samples per CUDA line). KEEP_REP=<workload> keeps that one .ncu-rep.
# TAG names the output directory under gpurun_out/. ~1 min of GPU time per workload.
set -u
TAG=${TAG:-r2_ncu_all}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
if [ -n "${PYTEST:-}" ]; then timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $OUT/pytest.txt; cat $OUT/pytest.txt; fi
for W in ${WORKLOADS:-xyzi c3 c4 lossless lz4 viz}; do
  timeout 600 ncu --set full --clock-control none --profile-from-start off --import-source on -f -o $OUT/tour_$W \
      python tools/kernel_tour.py --only $W --out $OUT/tour_$W.json > $OUT/tour_$W.log 2>&1
  echo "$W: exit $?"; tail -2 $OUT/tour_$W.log | cut -c1-300
  ncu -i $OUT/tour_$W.ncu-rep --page raw --csv > $OUT/tour_$W.raw.csv 2>/dev/null
  if [ -n "${SRC_KERNELS:-}" ]; then
    ncu -i $OUT/tour_$W.ncu-rep --page source --print-source cuda,sass --csv -k "regex:$SRC_KERNELS" > $OUT/tour_$W.source.csv 2>/dev/null
  fi
  [ "${KEEP_REP:-}" = "$W" ] || rm -f $OUT/tour_$W.ncu-rep
done
# the numbers outside the profiler, all workloads in one process
timeout 600 python tools/kernel_tour.py --out $OUT/tour_all.json > $OUT/tour_all.log 2>&1
du -sh $OUT; ls -la $OUT
