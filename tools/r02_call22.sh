N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --config5 --no-cpu-baseline > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"
tail -c 1500 gpurun_out/r02_bench_n$N.json; tail -5 gpurun_out/r02_bench_n$N.err
