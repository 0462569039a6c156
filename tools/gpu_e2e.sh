#!/usr/bin/env bash
# Development aid: parity suite + one bench line (e2e and extras included) with few steps.
set -u
TAG=${TAG:-r2_e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $OUT/pytest.txt; cat $OUT/pytest.txt
timeout 900 python bench.py --steps 30 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
python - <<PY
import json
b = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", json.dumps(b["e2e"])[:600])
print("dds", json.dumps(b.get("extras", {}).get("dds_converter_step"))[:700])
PY
