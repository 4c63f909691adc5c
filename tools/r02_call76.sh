set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for v in 1 4 8; do
  echo "== SEUNET_CAT_PPB=$v"
  SEUNET_CAT_PPB=$v timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_c76_layers_p$v.txt 2>&1; grep -E "cat:|total" gpurun_out/r02_c76_layers_p$v.txt
  SEUNET_CAT_PPB=$v timeout 300 python tools/time_forward.py 7 128 10 2>&1 | tail -1
done
( time timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py -m gpu -x -q ) > gpurun_out/r02_c76_tests_fwd.log 2>&1
tail -5 gpurun_out/r02_c76_tests_fwd.log
