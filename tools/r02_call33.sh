timeout 600 python -m pytest tests/test_gpu_wgrad.py tests/test_gpu_backward.py tests/test_gpu_training.py -q -x > gpurun_out/r02_tests_k.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_k.log
tail -5 gpurun_out/r02_tests_k.log
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_train_b8_v3.txt 2>&1; head -3 gpurun_out/r02_train_b8_v3.txt; grep wgrad: gpurun_out/r02_train_b8_v3.txt
SEUNET_WG_V3=0 timeout 300 python tools/time_train.py 8 128 2>&1 | head -2
timeout 300 python tools/time_train.py 1 128 2>&1 | head -2
