# window plans (head 0 skipped, head kernel sinks into the accumulator) + 256-bit gradient chunk IO
set -x
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c40_tests.log 2>&1; tail -3 gpurun_out/r02_c40_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_c40_bench.json 2> gpurun_out/r02_c40_bench.err; tail -c 300 gpurun_out/r02_c40_bench.err
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_c40_train_b8.txt 2>&1; head -3 gpurun_out/r02_c40_train_b8.txt
