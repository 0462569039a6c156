#!/usr/bin/env bash
# usage: tools/quick_bench.sh <label>   (env selects variants) — prints encode/decode ms per step of FRAMES (default 128) frames
timeout 300 python bench.py --frames ${FRAMES:-128} --steps 20 --warmup 3 --no-e2e --no-extras --cpu-seconds 0.2 > /tmp/vb.out 2>&1
tail -1 /tmp/vb.out > /tmp/vb.json
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open('/tmp/vb.json'))
except Exception:
    print("%-14s FAILED: %s" % (sys.argv[1], " | ".join(open('/tmp/vb.out').read().strip().splitlines()[-3:])[:400]))
    sys.exit(0)
print("%-14s enc_ms %.4f dec_ms %.4f enc_frac %.3f dec_frac %.3f parity %s value %.0f" % (sys.argv[1], d["config"]["encode_ms_per_step"], d["config"]["decode_ms_per_step"], d["roofline"]["encode"]["frac"], d["roofline"]["frac"], d["config"]["parity"], d["value"]))
PY
