mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train"
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1200 --csv --log-file gpurun_out/r02_c83_launches.csv $B > gpurun_out/r02_c83_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_shares.py gpurun_out/r02_c83_launches.csv > gpurun_out/r02_c83_shares.txt; head -30 gpurun_out/r02_c83_shares.txt
