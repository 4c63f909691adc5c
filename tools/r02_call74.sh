set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 300 python tools/layer_times.py 7 128 > gpurun_out/r02_c74_layers.txt 2>&1; grep -E "apply:(ec1|ec2|dc6)|total" gpurun_out/r02_c74_layers.txt
timeout 300 python tools/time_forward.py 7 128 10 2>&1 | tail -1
