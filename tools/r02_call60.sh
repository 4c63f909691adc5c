# full GPU suite on the new build (batched weight packing, head adjoint on a side stream, SSE pass B by recomputation),
# then A/B of the new knobs on the training step at B = 8 and B = 1, then the bench line
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c60_tests.log 2>&1
tail -5 gpurun_out/r02_c60_tests.log
for b in 8 1; do
  echo "== default B=$b"; DETAIL=1 timeout 300 python tools/time_train.py $b 128 2 > gpurun_out/r02_c60_train_b$b.txt 2>&1; head -3 gpurun_out/r02_c60_train_b$b.txt
  echo "== RECOMPUTE=0 B=$b"; SEUNET_BWD_RECOMPUTE=0 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -2
  echo "== HEAD_SIDE=0 B=$b"; SEUNET_BWD_HEAD_SIDE=0 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -1
  echo "== BWDA_FLOOR=2 B=$b"; SEUNET_BWDA_FLOOR=2 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -1
done
timeout 900 python bench.py > gpurun_out/r02_c60_bench.json 2> gpurun_out/r02_c60_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c60_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"])
PY
