set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for v in 1 4 8; do
  echo "== SEUNET_SSE_VPT=$v"
  SEUNET_SSE_VPT=$v timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_c69_layers_vpt$v.txt 2>&1; grep -E "apply:(ec3|ec4|ec5|ec6|ec7|ec9|ec10|dc1|dc3|dc4|dc5)|total" gpurun_out/r02_c69_layers_vpt$v.txt
done
for v in 1 4 8; do SEUNET_SSE_VPT=$v timeout 300 python tools/time_train.py 8 128 2 2>&1 | head -1; done
for v in 1 4 8; do SEUNET_SSE_VPT=$v timeout 300 python tools/time_forward.py 7 128 10 2>&1 | tail -1; done
