timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c59_bench_n8.json 2> gpurun_out/r02_c59_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/r02_c59_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_c59_bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']); print(d.get('train',{}).get('ms_per_step'), d.get('train',{}).get('value')); print(d.get('sweep'))"
