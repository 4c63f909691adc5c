export CONV_BENCH_ITERS=1 CONV_BENCH_WARMUP=0
python tools/conv_layer_bench.py 7 128 ec3 > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 1 -f -o gpurun_out/r02_conv_ec3 python tools/conv_layer_bench.py 7 128 ec3 > gpurun_out/r02_ncu_ec3.log 2>&1
tail -3 gpurun_out/r02_ncu_ec3.log
