timeout 1200 python -m pytest tests/test_gpu_multi_device.py -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02_c54_bench_n2.json 2> gpurun_out/r02_c54_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_c54_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_c54_bench_n2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
