nvidia-smi -L
python -m pytest tests/test_gpu_multi_device.py -q -s > gpurun_out/r02_tests_mg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_mg.log
tail -40 gpurun_out/r02_tests_mg.log
