timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py -q -x > gpurun_out/r02_tests_c.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_c.log
tail -4 gpurun_out/r02_tests_c.log
CONV_BENCH_ITERS=2 CONV_BENCH_WARMUP=1 SEUNET_LIB_PATH=tools/libseunet_prof.so python tools/conv_layer_bench.py 7 128 dc5,dc6,ec2,ec3,dc3,ec4 2>&1 | awk '/conv prof/{l=$0} !/conv prof/{print l; print $0}' > gpurun_out/r02_convprof2.txt
cat gpurun_out/r02_convprof2.txt
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v2.txt 2>&1
grep -E "conv:|total" gpurun_out/r02_layers_v2.txt
