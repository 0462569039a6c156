#!/usr/bin/env bash
# One GPU call for the round's evidence: parity suite, bench line (with extras), config 4 (1000 frames, encode + LZ4 on the
# device), launch lists. TAG names the output directory under gpurun_out/.
set -u
TAG=${TAG:-r2_round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $OUT/pytest.txt
cat $OUT/pytest.txt
timeout 900 python bench.py ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err
tail -c 1500 $OUT/bench.json
timeout 600 python bench.py --frames 32 --steps 100 --warmup 5 --no-e2e --no-extras > $OUT/bench_f32.json 2> $OUT/bench_f32.err
tail -c 600 $OUT/bench_f32.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
tail -c 800 $OUT/bench_reference.json
timeout 900 python tools/config4_bench.py --frames ${C4_FRAMES:-1000} > $OUT/config4.json 2> $OUT/config4.err
cat $OUT/config4.json; tail -3 $OUT/config4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > $OUT/launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_extras.csv \
    python tools/extras_bench.py > $OUT/launches_extras.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_config4.csv \
    python tools/config4_bench.py --frames 200 --reps 2 > $OUT/launches_config4.log 2>&1
ls -la $OUT
