timeout 600 python -m pytest tests/test_gpu_forward.py -q -x 2>&1 | tail -2
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_layers_v9.txt 2>&1
grep -E "apply:ec1|apply:ec2|apply:dc6|total" gpurun_out/r02_layers_v9.txt
