set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train"
$B > gpurun_out/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1500 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
F="python tools/layer_times.py 7 128"
$F > gpurun_out/r02_layers_v7.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 44 -c 22 -o gpurun_out/r02_conv_b7 $F > gpurun_out/r02_ncu_conv.log 2>&1
echo "conv full rc=$?"
T="python tools/time_train.py 8 128"
$T > gpurun_out/r02_train_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 48 -c 24 -o gpurun_out/r02_wgrad_b8 $T > gpurun_out/r02_ncu_wgrad.log 2>&1
echo "wgrad full rc=$?"
tail -1 gpurun_out/r02_layers_v7.txt
ls -la gpurun_out/*.ncu-rep
