timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py tests/test_gpu_parity_full.py -q -x > gpurun_out/r02_tests_h.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_h.log
tail -4 gpurun_out/r02_tests_h.log
SEUNET_CAT_FUSION=0 timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v5_unfused.txt 2>&1
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v5_fused.txt 2>&1
grep -E "ec3 |ec33|ec6 |ec63|dc4 |dc42|ec9 |ec93|ec12 |ec123|dc2 |dc22|total" gpurun_out/r02_layers_v5_unfused.txt
echo ---
grep -E "ec3 |ec33|ec6 |ec63|dc4 |dc42|ec9 |ec93|ec12 |ec123|dc2 |dc22|total" gpurun_out/r02_layers_v5_fused.txt
timeout 300 python tools/layer_times.py 1 128 | tail -1
