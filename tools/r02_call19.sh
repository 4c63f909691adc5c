python -m pytest tests/test_gpu_multi_device.py -q > gpurun_out/r02_tests_mg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_mg.log; tail -3 gpurun_out/r02_tests_mg.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --config5 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
for n in (1,2):
    try:
        j=json.loads(open(f'gpurun_out/r02_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'ms', round(j['ms_per_step'],2), 'e2e ms', round(j['e2e']['ms_per_step'],2), 'pp', j.get('e2e_postprocessed',{}).get('ms_per_step'), 'sweep', (j.get('sweep') or {}).get('ms_per_volume_per_gpu'), 'train', (j.get('train') or {}).get('ms_per_step'), (j.get('train') or {}).get('cpu_baseline'), 'train5', (j.get('train_s160') or {}).get('ms_per_step'), j['roofline']['frac'], j['roofline']['whole_step_frac'])
    except Exception as e: print(n, 'ERR', e, open(f'gpurun_out/r02_bench_n{n}.err').read()[-1500:])
PY
