timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r02_tests_g.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_g.log
tail -4 gpurun_out/r02_tests_g.log
timeout 600 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_a.json
