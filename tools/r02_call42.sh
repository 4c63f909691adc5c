# one-pass trunk up-sampling adjoint + 32-bit/float4 head adjoint: correctness, then timing against the three-pass version on the same box
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -x -q 2>&1 | tail -3
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_c42_train_b8.txt 2>&1; head -2 gpurun_out/r02_c42_train_b8.txt; grep -E "bwd:(head|up)" gpurun_out/r02_c42_train_b8.txt
SEUNET_UPBWD_3PASS=1 DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_c42_train_b8_3pass.txt 2>&1; head -2 gpurun_out/r02_c42_train_b8_3pass.txt; grep -E "bwd:(head|up)" gpurun_out/r02_c42_train_b8_3pass.txt
DETAIL=1 timeout 300 python tools/time_train.py 1 128 > gpurun_out/r02_c42_train_b1.txt 2>&1; head -2 gpurun_out/r02_c42_train_b1.txt; grep -E "bwd:(head|up)" gpurun_out/r02_c42_train_b1.txt
