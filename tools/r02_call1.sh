set -x
for g in ss layout sw128 ts m64 cp mix cg2; do timeout 120 tools/umma_bench $g > gpurun_out/r02_umma_$g.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_umma_$g.txt; done
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/r02_clk_layers.csv &
SMI=$!
python tools/layer_times.py 7 128 > gpurun_out/r02_layers_base.txt 2>&1
CONV_BENCH_ITERS=200 python tools/conv_layer_bench.py 7 128 dc5 > gpurun_out/r02_dc5_loop.txt 2>&1
CONV_BENCH_ITERS=200 python tools/conv_layer_bench.py 7 128 dc3 >> gpurun_out/r02_dc5_loop.txt 2>&1
kill $SMI
cat gpurun_out/r02_umma_*.txt
