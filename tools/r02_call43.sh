timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -x -q 2>&1 | tail -3
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_c43_train_b8.txt 2>&1; head -2 gpurun_out/r02_c43_train_b8.txt; grep -E "bwd:(head|up)|bwdA:(ec33|ec63|ec93|dc5)" gpurun_out/r02_c43_train_b8.txt
T="python tools/time_train.py 8 128"
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/c43_$1 $T > /dev/null 2>&1
  ncu -i /tmp/c43_$1.ncu-rep --page details > gpurun_out/r02_c43_$1.details.txt 2>/dev/null
  ncu -i /tmp/c43_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_c43_$1.sass.csv 2>/dev/null
}
cap catbwd_ec33 cat_bwd_a_kernel 5 1
cap ssebwd_dc5 sse_bwd_a_kernel 1 1
cap upbwd_d1 upsample2_bwd_fused_kernel 0 1
cap normb_dc5 norm_bwd_b_kernel 1 1
