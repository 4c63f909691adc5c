# sse_bwd_a ring v2 (elected-lane bulk copies, 1-2 voxel groups per slot): parity, A/B, ncu
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_knobs.py -m gpu -x -q ) > gpurun_out/r02_c64_tests_bwd.log 2>&1
tail -5 gpurun_out/r02_c64_tests_bwd.log
for b in 8 1; do
  echo "== default (ring) B=$b"; DETAIL=1 timeout 300 python tools/time_train.py $b 128 2 > gpurun_out/r02_c64_train_b$b.txt 2>&1; head -3 gpurun_out/r02_c64_train_b$b.txt
  echo "== RING=0 B=$b"; SEUNET_BWDA_RING=0 timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -2
done
T="python tools/time_train.py 8 128"
ncu --set full --clock-control none --import-source on -k regex:sse_bwd_a_kernel -s 1 -c 1 -o /tmp/c64_ring $T > /dev/null 2>&1
ncu -i /tmp/c64_ring.ncu-rep --page details > gpurun_out/r02_c64_ring.details.txt 2>/dev/null
ncu -i /tmp/c64_ring.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_c64_ring.sass.csv 2>/dev/null
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Executed Ipc Active" gpurun_out/r02_c64_ring.details.txt
