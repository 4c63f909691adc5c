# pass A of the SSE blocks with a bulk-copy (cp.async.bulk + mbarrier) input ring: parity first, then A/B
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_knobs.py -m gpu -x -q ) > gpurun_out/r02_c62_tests_bwd.log 2>&1
tail -5 gpurun_out/r02_c62_tests_bwd.log
for b in 8 1; do
  echo "== default (ring) B=$b"; DETAIL=1 timeout 300 python tools/time_train.py $b 128 2 > gpurun_out/r02_c62_train_b$b.txt 2>&1; head -3 gpurun_out/r02_c62_train_b$b.txt
  echo "== RING=0 B=$b"; SEUNET_BWDA_RING=0 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -2
  echo "== ring + RECOMPUTE=1 B=$b"; SEUNET_BWD_RECOMPUTE=1 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -2
done
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c62_tests.log 2>&1
tail -5 gpurun_out/r02_c62_tests.log
