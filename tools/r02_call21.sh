timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py tests/test_gpu_backward.py -q -x > gpurun_out/r02_tests_e.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_e.log
tail -3 gpurun_out/r02_tests_e.log
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v3.txt 2>&1
grep -E "conv:(ec1|ec2|ec3|ec4|ec6|dc3|dc4|dc5|dc6)|total" gpurun_out/r02_layers_v3.txt
