timeout 900 python -m pytest tests/test_gpu_wgrad.py tests/test_gpu_backward.py tests/test_gpu_training.py -q -x > gpurun_out/r02_tests_i.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_i.log
tail -5 gpurun_out/r02_tests_i.log
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_train_b8_kh.txt 2>&1; head -4 gpurun_out/r02_train_b8_kh.txt; grep wgrad: gpurun_out/r02_train_b8_kh.txt
SEUNET_WG_NKH=1 timeout 300 python tools/time_train.py 8 128 2>&1 | head -3
