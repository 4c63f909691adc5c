set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for v in 4 8; do
  echo "== SEUNET_SSE_VPT_NARROW=$v"
  SEUNET_SSE_VPT_NARROW=$v timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_c70_layers_n$v.txt 2>&1; grep -E "apply:(ec1|ec2|dc6)|total" gpurun_out/r02_c70_layers_n$v.txt
  SEUNET_SSE_VPT_NARROW=$v timeout 300 python tools/time_forward.py 7 128 10 2>&1 | tail -1
done
( time timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py tests/test_gpu_backward.py -m gpu -x -q ) 2>&1 | tail -4
