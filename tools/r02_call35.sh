timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r02_tests_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_final.log
tail -3 gpurun_out/r02_tests_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_final.txt 2>&1; tail -1 gpurun_out/r02_layers_final.txt
DETAIL=1 timeout 300 python tools/time_train.py 8 128 > gpurun_out/r02_train_b8_final.txt 2>&1; head -3 gpurun_out/r02_train_b8_final.txt
DETAIL=1 timeout 300 python tools/time_train.py 1 128 > gpurun_out/r02_train_b1_final.txt 2>&1; head -3 gpurun_out/r02_train_b1_final.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; echo "ref rc=$?"; tail -c 600 gpurun_out/r02_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
