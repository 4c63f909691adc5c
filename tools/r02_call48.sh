timeout 1200 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py tests/test_gpu_conv.py -x -q 2>&1 | tail -3
python tools/write_bw.py
echo "== old kernels"; SEUNET_UP_WARP=0 SEUNET_PREP_VEC=0 python tools/layer_times.py 7 128 | grep -E "^(prep|up:|total)"
echo "== new kernels"; python tools/layer_times.py 7 128 | grep -E "^(prep|up:|total)"
for cfg in "7 3" "14 3" "14 2" "21 2" "14 1" "7 1"; do set -- $cfg; echo "== batch $1 streams $2"; timeout 300 python bench.py --no-cpu-baseline --no-train --batch $1 --streams $2 --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"; done
