# shallow dgrad tiles (A/B), recompute default off, averaged per-layer timing in the bench; full suite, then the bench line
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c61_tests.log 2>&1
tail -5 gpurun_out/r02_c61_tests.log
for b in 1 2 8; do
  echo "== default B=$b"; DETAIL=1 timeout 300 python tools/time_train.py $b 128 2 > gpurun_out/r02_c61_train_b$b.txt 2>&1; head -3 gpurun_out/r02_c61_train_b$b.txt
  echo "== SHALLOW=0 B=$b"; SEUNET_CONV_SHALLOW=0 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -3
done
timeout 900 python bench.py > gpurun_out/r02_c61_bench.json 2> gpurun_out/r02_c61_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c61_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"], d["roofline"]["achieved"], d["roofline"]["conv_ms_per_forward"], d["roofline"]["other_ms_per_forward"])
PY
