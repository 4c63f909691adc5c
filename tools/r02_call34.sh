timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -q -x > gpurun_out/r02_tests_l.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_l.log
tail -3 gpurun_out/r02_tests_l.log
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_train_b8_v4.txt 2>&1; head -3 gpurun_out/r02_train_b8_v4.txt; grep -E "bwdA:ec33|bwdA:ec63|bwdA:ec93" gpurun_out/r02_train_b8_v4.txt
