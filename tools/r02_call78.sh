set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_weight_cache.py -m gpu -x -q ) > gpurun_out/r02_c78_tests.log 2>&1
tail -25 gpurun_out/r02_c78_tests.log
