python -m pytest tests -m gpu -q -s --maxfail=40 --deselect tests/test_gpu_multi_device.py > gpurun_out/r02_tests_a.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_a.log
tail -5 gpurun_out/r02_tests_a.log
grep -E "trajectory|after 40|512x512x400|weights x30|largest stored|output-gradient|tightest|vs plain|FAILED|Error|error" gpurun_out/r02_tests_a.log | head -60
