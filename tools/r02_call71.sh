set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c71_tests.log 2>&1
tail -6 gpurun_out/r02_c71_tests.log
timeout 900 python bench.py > gpurun_out/r02_c71_bench.json 2> gpurun_out/r02_c71_bench.err
echo "bench rc=$?"
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_c71_layers.txt 2>&1; tail -1 gpurun_out/r02_c71_layers.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c71_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"], d["roofline"]["achieved"], d["roofline"]["conv_ms_per_forward"], d["roofline"]["other_ms_per_forward"])
PY
