#!/usr/bin/env bash
# First gpurun of a round: everything that was written / changed without a GPU gets its hardware run and numbers.
#   gpurun --timeout 1500 -- 'bash tools_round_start_gpu.sh'
# Results land in gpurun_out/round_start/ (copy what matters into profiles/).
set -u
OUT=gpurun_out/round_start
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1
echo "== pytest -m gpu" | tee "$OUT/summary.txt"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee -a "$OUT/summary.txt"
echo "== random-layout sweep through the generic kernels" | tee -a "$OUT/summary.txt"
CLDN_B200_FUZZ=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k random_layouts_sweep 2>&1 | tail -3 | tee -a "$OUT/summary.txt"
echo "== corrupted blobs, more seeds" | tee -a "$OUT/summary.txt"
for s in 1 2 3; do CLDN_B200_CORRUPT_SEED=$s CLDN_B200_CORRUPT_TRIALS=100 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k corrupted 2>&1 | tail -1 | tee -a "$OUT/summary.txt"; done
echo "== differential fuzzers on the hardware (tests/fuzz), new kernels selected" | tee -a "$OUT/summary.txt"
export CLDN_B200_UNMEASURED=1
timeout 600 python tests/fuzz/fuzz_corrupt_blobs.py 1 150 2>&1 | tail -15 | tee -a "$OUT/summary.txt"
timeout 300 python tests/fuzz/fuzz_gorilla_records.py 1 500 2>&1 | tail -2 | tee -a "$OUT/summary.txt"
timeout 300 python tests/fuzz/fuzz_encode_inputs.py 1 100 2>&1 | tail -2 | tee -a "$OUT/summary.txt"
unset CLDN_B200_UNMEASURED
echo "== bench.py (N=1)" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 3000 "$OUT/bench_n1.json" | tee -a "$OUT/summary.txt"
echo "== config table / decode-path A-B" | tee -a "$OUT/summary.txt"
timeout 600 python tools_config_bench.py > "$OUT/config_table.jsonl" 2>&1; cat "$OUT/config_table.jsonl" | tee -a "$OUT/summary.txt"
timeout 900 python tools_ab_decode_paths.py > "$OUT/ab_decode_paths.jsonl" 2>&1; cat "$OUT/ab_decode_paths.jsonl" | tee -a "$OUT/summary.txt"
echo "== extras with the default kernels and with CLDN_B200_UNMEASURED=1 (parallel run-table reader in C3, ...)" | tee -a "$OUT/summary.txt"
timeout 600 python tools_extras_bench.py 2>&1 | tail -1 | tee "$OUT/extras_default.json" | tee -a "$OUT/summary.txt"
CLDN_B200_UNMEASURED=1 timeout 600 python tools_extras_bench.py 2>&1 | tail -1 | tee "$OUT/extras_unmeasured.json" | tee -a "$OUT/summary.txt"
echo "== ncu launch list of the extras (new kernels)" | tee -a "$OUT/summary.txt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/extras_launches.csv" python tools_extras_bench.py > "$OUT/extras_under_ncu.log" 2>&1
python - <<'PY' | tee -a gpurun_out/round_start/summary.txt
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/round_start/extras_launches.csv", errors="replace")) if len(r) > 5]
hdr = next((r for r in rows if "Kernel Name" in r), None)
if hdr:
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) <= max(k, v):
            continue
        try:
            agg[r[k][:70]][0] += 1; agg[r[k][:70]][1] += float(r[v].replace(",", ""))
        except ValueError:
            pass
    for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:20]:
        print(f"{t / 1e3:10.1f} us total {n:5d} launches  {name}")
PY
echo "== ncu --set full: viz kernels, decode_mixed, decode_gorilla (one launch each)" | tee -a "$OUT/summary.txt"
for k in viz_insert_kernel viz_compact_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 1 -o "$OUT/full_$k" -f python tools_extras_bench.py > "$OUT/full_$k.log" 2>&1 || true
done
for k in decode_mixed_kernel decode_gorilla_kernel gorilla_prepass_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 1 -o "$OUT/full_$k" -f python tools_ab_decode_paths.py > "$OUT/full_$k.log" 2>&1 || true
done
ls -la "$OUT" | tee -a "$OUT/summary.txt"
