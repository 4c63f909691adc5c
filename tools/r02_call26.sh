timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py tests/test_gpu_parity_full.py -q -x > gpurun_out/r02_tests_h.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_h.log
tail -4 gpurun_out/r02_tests_h.log
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v6.txt 2>&1
grep -E "ec3 |ec33|dc4 |dc42|total" gpurun_out/r02_layers_v6.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_sse_cat -s 2 -c 1 -o gpurun_out/r02_catfuse python tools/layer_times.py 7 128 > gpurun_out/r02_ncu_catfuse.log 2>&1; echo "ncu rc=$?"
