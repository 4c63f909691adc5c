set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train"
$B > gpurun_out/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1200 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
F="python tools/layer_times.py 7 128"
$F > gpurun_out/r02_layers_v7.txt 2>&1 && ncu --set full --clock-control none -k regex:conv_tc_kernel -s 44 -c 22 -o /tmp/r02_conv_b7 $F > gpurun_out/r02_ncu_conv.log 2>&1
echo "conv full rc=$?"
ncu -i /tmp/r02_conv_b7.ncu-rep --page raw --csv > gpurun_out/r02_conv_b7_raw.csv 2>/dev/null
T="python tools/time_train.py 8 128"
$T > gpurun_out/r02_train_plain.log 2>&1 && ncu --set full --clock-control none -k regex:wgrad_tc_kernel -s 48 -c 24 -o /tmp/r02_wgrad_b8 $T > gpurun_out/r02_ncu_wgrad.log 2>&1
echo "wgrad full rc=$?"
ncu -i /tmp/r02_wgrad_b8.ncu-rep --page raw --csv > gpurun_out/r02_wgrad_b8_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:sse_bwd_a -s 39 -c 1 -o gpurun_out/r02_bwda_dc3 $T > gpurun_out/r02_ncu_bwda.log 2>&1
echo "bwda rc=$?"
tail -1 gpurun_out/r02_layers_v7.txt; grep head gpurun_out/r02_layers_v7.txt
ls -la gpurun_out/ | tail -12
