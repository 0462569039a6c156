#!/usr/bin/env bash
# Development aid: build, GPU parity suite, then the event-timed legs of tools/kernel_tour.py for the workloads in WORKLOADS.
set -u
TAG=${TAG:-r2_quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $OUT/pytest.txt; cat $OUT/pytest.txt
timeout 900 python tools/kernel_tour.py --only ${WORKLOADS:-c3,c4} --out $OUT/tour.json > $OUT/tour.log 2>&1; tail -c 3000 $OUT/tour.log
if [ -n "${LAUNCHES:-}" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv \
      python tools/kernel_tour.py --only ${WORKLOADS:-c3,c4} --out $OUT/tour_ncu.json > $OUT/launches.log 2>&1
fi
