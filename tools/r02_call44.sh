timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -x -q 2>&1 | tail -3
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_c44_train_b8.txt 2>&1; head -2 gpurun_out/r02_c44_train_b8.txt; grep -E "bwd:(head|up)|bwdA:(ec33|ec63|ec93|dc5)" gpurun_out/r02_c44_train_b8.txt
DETAIL=1 timeout 300 python tools/time_train.py 1 128 > gpurun_out/r02_c44_train_b1.txt 2>&1; head -2 gpurun_out/r02_c44_train_b1.txt; grep -E "bwd:(head|up)|bwdA:(ec33|ec63|ec93|dc5)" gpurun_out/r02_c44_train_b1.txt
