# state check at HEAD after the container was re-created: full GPU suite, full bench line, then ncu of the backward HBM-bound kernels
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c39_tests.log 2>&1; tail -3 gpurun_out/r02_c39_tests.log
timeout 900 python bench.py > gpurun_out/r02_c39_bench.json 2> gpurun_out/r02_c39_bench.err; tail -c 600 gpurun_out/r02_c39_bench.json
T="python tools/time_train.py 2 128"
DETAIL=1 $T > gpurun_out/r02_c39_train_b2.txt 2>&1 || exit 1
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/c39_$1 $T > /dev/null 2>&1
  ncu -i /tmp/c39_$1.ncu-rep --page details > gpurun_out/r02_c39_$1.details.txt 2>/dev/null
  ncu -i /tmp/c39_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_c39_$1.sass.csv 2>/dev/null
}
cap catbwd_ec33 cat_bwd_a_kernel 5 1
cap ssebwd_dc5 sse_bwd_a_kernel 1 1
cap upbwd_d1 upsample2_bwd_kernel 0 1
cap adjoint adjoint_axis 0 8
cap catbwdx cat_bwd_x_kernel 2 1
ls -la gpurun_out/
