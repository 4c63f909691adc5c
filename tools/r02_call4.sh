timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py tests/test_gpu_backward.py -q -x > gpurun_out/r02_tests_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_b.log
tail -15 gpurun_out/r02_tests_b.log
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v1.txt 2>&1
grep -E "conv:|total" gpurun_out/r02_layers_v1.txt
