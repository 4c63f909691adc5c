timeout 600 python -m pytest tests/test_gpu_wgrad.py tests/test_gpu_backward.py tests/test_gpu_training.py -q -x 2>&1 | tail -2
DETAIL=1 timeout 300 python tools/time_train.py 1 128 > gpurun_out/r02_train_b1_seg.txt 2>&1; head -2 gpurun_out/r02_train_b1_seg.txt; grep -E "wgrad:(ec1|ec2|ec3|ec4|ec5|dc4|dc5|dc6)" gpurun_out/r02_train_b1_seg.txt
timeout 300 python tools/time_train.py 8 128 2>&1 | head -2
timeout 300 python tools/time_train.py 2 128 2>&1 | head -2
