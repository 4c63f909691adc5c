set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_knobs.py -m gpu -q ) > gpurun_out/r02_c68_tests_bwd.log 2>&1
tail -30 gpurun_out/r02_c68_tests_bwd.log
( compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_backward.py::test_backward_matches_oracle_autograd_at_same_forward_state" -m gpu -q -k "24-24-40 or shape3" ) > gpurun_out/r02_c68_memcheck.log 2>&1
tail -15 gpurun_out/r02_c68_memcheck.log
