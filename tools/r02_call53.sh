timeout 1200 python -m pytest tests/test_gpu_multi_device.py -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_c53_bench_n2.json 2> gpurun_out/r02_c53_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02_c53_bench_n2.json
