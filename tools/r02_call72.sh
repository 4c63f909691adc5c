set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for v in 1 4; do
  echo "== SEUNET_BWDB_VPT=$v"
  for b in 8 1; do SEUNET_BWDB_VPT=$v timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -2; done
done
( time timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_knobs.py -m gpu -x -q ) > gpurun_out/r02_c72_tests_bwd.log 2>&1
tail -5 gpurun_out/r02_c72_tests_bwd.log
