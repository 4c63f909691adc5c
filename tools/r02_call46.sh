set -o pipefail
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r02_c46_bench.json 2> gpurun_out/r02_c46_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_c46_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c46_bench_ref.json 2>/dev/null; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
