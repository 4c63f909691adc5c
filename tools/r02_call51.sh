timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -40
echo "=== with old forward kernels"
SEUNET_SSE_SPLIT=0 SEUNET_UP_WARP=0 SEUNET_PREP_VEC=0 timeout 900 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -5
