#!/usr/bin/env bash
# Times the encode / decode legs of bench.py for the FloatN encode kernel variants (development helper).
for v in 0 1 2; do
  CLDN_B200_ENC_VARIANT=$v timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-extras --cpu-seconds 0.2 2>&1 | tail -1 > /tmp/vb.json
  python - <<PY
import json
d = json.load(open('/tmp/vb.json'))
print("variant $v enc_ms %.4f dec_ms %.4f enc_frac %.3f dec_frac %.3f parity %s" % (d["config"]["encode_ms_per_step"], d["config"]["decode_ms_per_step"], d["roofline"]["encode"]["frac"], d["roofline"]["frac"], d["config"]["parity"]))
PY
done
