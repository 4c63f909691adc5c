# final-build verification (after the per-block-prologue changes): full GPU suite, bench line, training breakdowns, ncu launch list of the bench command
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c80_tests.log 2>&1
tail -5 gpurun_out/r02_c80_tests.log
timeout 900 python bench.py > gpurun_out/r02_c80_bench.json 2> gpurun_out/r02_c80_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_c80_bench_ref.json 2> gpurun_out/r02_c80_bench_ref.err
echo "ref rc=$?"
for b in 8 1; do DETAIL=1 timeout 300 python tools/time_train.py $b 128 2 > gpurun_out/r02_c80_train_b$b.txt 2>&1; head -3 gpurun_out/r02_c80_train_b$b.txt; done
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_c80_layers.txt 2>&1; tail -2 gpurun_out/r02_c80_layers.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c80_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"], d["roofline"]["achieved"], d["roofline"]["conv_ms_per_forward"], d["roofline"]["other_ms_per_forward"])
PY
