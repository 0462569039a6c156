set -u
OUT=gpurun_out/r2_start
mkdir -p $OUT
{
echo "== nvcomp / lz4 / zstd probe"
find / -xdev \( -iname '*nvcomp*' -o -name 'liblz4*' -o -name 'libzstd*' \) 2>/dev/null | grep -v '^/proc' | head -40
python -c "import importlib.util as u; print('nvidia.nvcomp', u.find_spec('nvidia.nvcomp') if u.find_spec('nvidia') else None)" 2>&1
pip list 2>/dev/null | grep -i -E 'nvcomp|kvikio|cupy|lz4|zstd' 
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
lscpu | grep -E 'Model name|^CPU\(s\)|NUMA'
nvidia-smi topo -m | head -20
} > $OUT/probe.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest gpu default" > $OUT/summary.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 >> $OUT/summary.txt
echo "== pytest gpu UNMEASURED=1" >> $OUT/summary.txt
CLDN_B200_UNMEASURED=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 >> $OUT/summary.txt
echo "== sweep" >> $OUT/summary.txt
CLDN_B200_FUZZ=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k random_layouts_sweep 2>&1 | tail -3 >> $OUT/summary.txt
echo "== extras" >> $OUT/summary.txt
timeout 400 python tools_extras_bench.py 2>&1 | tail -1 > $OUT/extras_default.json
CLDN_B200_UNMEASURED=1 timeout 400 python tools_extras_bench.py 2>&1 | tail -1 > $OUT/extras_unmeasured.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/extras_launches.csv python tools_extras_bench.py > $OUT/extras_under_ncu.log 2>&1
CLDN_B200_UNMEASURED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/extras_launches_unmeasured.csv python tools_extras_bench.py > $OUT/extras_under_ncu2.log 2>&1
cat $OUT/summary.txt
