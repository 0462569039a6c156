#!/usr/bin/env bash
# tests + extras + a sweep of the bench's batch size (frames per step)
set -u
TAG=${TAG:-r2_sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $OUT/pytest.txt
cat $OUT/pytest.txt
timeout 600 python tools/extras_bench.py 2>&1 | tail -1 > $OUT/extras.json
cat $OUT/extras.json | cut -c1-3000
for F in ${FRAMES:-32 64 100 128}; do
  timeout 600 python bench.py --frames $F --steps 30 --warmup 5 --no-e2e --no-extras 2>/dev/null | tail -1 > $OUT/bench_F$F.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_F$F.json"))
print("F=$F value", round(d["value"]), "enc_ms", round(d["config"]["encode_ms_per_step"],4), "dec_ms", round(d["config"]["decode_ms_per_step"],4), "enc_frac", round(d["roofline"]["encode"]["frac"],3), "dec_frac", round(d["roofline"]["frac"],3))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_extras.csv \
    python tools/extras_bench.py > $OUT/launches_extras.log 2>&1
