timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r02_tests_j.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_j.log
tail -4 gpurun_out/r02_tests_j.log
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v8.txt 2>&1
grep -E "prep|ec1 |up:|head|total" gpurun_out/r02_layers_v8.txt
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_train_b8_v2.txt 2>&1; head -3 gpurun_out/r02_train_b8_v2.txt
