python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_multi_device.py > gpurun_out/r02_tests_d.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_d.log
tail -8 gpurun_out/r02_tests_d.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_bench_b7.json 2> gpurun_out/r02_bench_b7.err; tail -c 600 gpurun_out/r02_bench_b7.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --batch 8 > gpurun_out/r02_bench_b8.json 2> gpurun_out/r02_bench_b8.err
python - <<'PY'
import json
for b in (7,8):
    try:
        j=json.loads(open(f'gpurun_out/r02_bench_b{b}.json').read().strip().splitlines()[-1])
        print(b, j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['conv_ms_per_forward'], j['roofline']['other_ms_per_forward'], j['clocks'])
    except Exception as e: print(b, 'ERR', e, open(f'gpurun_out/r02_bench_b{b}.err').read()[-800:])
PY
