timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py tests/test_gpu_parity_full.py tests/test_gpu_wgrad.py tests/test_gpu_backward.py tests/test_gpu_training.py -q -x > gpurun_out/r02_tests_f.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_f.log
tail -6 gpurun_out/r02_tests_f.log
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v4.txt 2>&1
grep -E "ec3|ec33|ec6|ec63|dc4|dc42|total" gpurun_out/r02_layers_v4.txt
timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_train_b8.txt 2>&1; tail -4 gpurun_out/r02_train_b8.txt
timeout 300 python tools/time_train.py 1 128 > gpurun_out/r02_train_b1.txt 2>&1; tail -4 gpurun_out/r02_train_b1.txt
