#!/usr/bin/env bash
# The differential fuzzers of tests/fuzz against the PRODUCT library on the B200 (they compare with oracle/_ref), plus smoke().
set -u
TAG=${TAG:-r2_fuzz}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
( timeout 600 python tests/fuzz/fuzz_v5_int_fields.py ${SEED:-101} 400 2>&1 | tail -2 ) | tee $OUT/v5_int_fields.txt
( timeout 600 python tests/fuzz/fuzz_corrupt_blobs.py ${SEED:-7} 1500 2>&1 | tail -4 ) | tee $OUT/corrupt_blobs.txt
( timeout 600 python tests/fuzz/fuzz_point_steps.py ${SEED:-5} 400 2>&1 | tail -2 ) | tee $OUT/point_steps.txt
( timeout 600 python tests/fuzz/fuzz_encode_inputs.py ${SEED:-5} 200 2>&1 | tail -2 ) | tee $OUT/encode_inputs.txt
( timeout 600 python tests/fuzz/fuzz_dds_messages.py ${SEED:-5} 300 2>&1 | tail -2 ) | tee $OUT/dds_messages.txt
