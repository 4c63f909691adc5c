# 2 GPUs, final build: multi-device tests (nn.DataParallel replicas, 2-rank trainer step, patch-sharded inference), bench at N = 2
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
( time timeout 1200 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_weight_cache.py -m gpu -x -q ) > gpurun_out/r02_c66_tests_md.log 2>&1
tail -4 gpurun_out/r02_c66_tests_md.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/r02_c66_bench_n2.json 2> gpurun_out/r02_c66_bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c66_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"])
PY
