#!/usr/bin/env bash
# One GPU call: parity suite, bench line, launch list and one full ncu capture of the kernels named in $KERNELS.
#   gpurun --timeout 1200 -- 'KERNELS="decode_floatn_fast" TAG=r2_dec1 bash tools/gpu_decode_check.sh'
set -u
TAG=${TAG:-r2_x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $OUT/pytest.txt
cat $OUT/pytest.txt
timeout 900 python bench.py ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err
tail -c 2500 $OUT/bench.json
for k in ${KERNELS:-}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -o $OUT/full_$k -f \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-extras > $OUT/ncu_$k.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > $OUT/launches.log 2>&1
ls -la $OUT
