set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tools/dp_check.py 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 > gpurun_out/r02_c79_bench_n2.json 2> gpurun_out/r02_c79_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_c79_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c79_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"])
PY
