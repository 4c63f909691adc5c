timeout 1200 python -m pytest tests/test_gpu_forward.py tests/test_gpu_sliding_window.py -x -q 2>&1 | tail -3
F='^(apply:(ec4|ec5|ec6|ec7|ec9|ec10|dc1|dc3|dc5|ec2|dc6)|total)'
echo "== split off"; SEUNET_SSE_SPLIT=0 python tools/layer_times.py 7 128 | grep -E "$F"
echo "== split 64"; SEUNET_SSE_SPLIT=1 python tools/layer_times.py 7 128 | grep -E "$F"
echo "== split 64+32"; SEUNET_SSE_SPLIT=3 python tools/layer_times.py 7 128 | grep -E "$F"
