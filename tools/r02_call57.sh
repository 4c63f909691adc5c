timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02_c57_bench.json 2> gpurun_out/r02_c57_bench.err; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1200 --csv --log-file gpurun_out/r02_c57_launches.csv $B > gpurun_out/r02_c57_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_shares.py gpurun_out/r02_c57_launches.csv > gpurun_out/r02_c57_shares.txt; cat gpurun_out/r02_c57_shares.txt
python tools/layer_times.py 7 128 > gpurun_out/r02_c57_layers.txt; tail -1 gpurun_out/r02_c57_layers.txt
DETAIL=1 python tools/time_train.py 8 128 > gpurun_out/r02_c57_train_b8.txt 2>&1; head -2 gpurun_out/r02_c57_train_b8.txt
DETAIL=1 python tools/time_train.py 1 128 > gpurun_out/r02_c57_train_b1.txt 2>&1; head -2 gpurun_out/r02_c57_train_b1.txt
ls -la gpurun_out/r02_c57_launches.csv
