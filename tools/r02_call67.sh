# N-GPU bench line of the final build (usage: bash tools/r02_call67.sh N)
N=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N > gpurun_out/r02_c67_bench_n$N.json 2> gpurun_out/r02_c67_bench_n$N.err
echo "bench n$N rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_c67_bench_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["train"]["ms_per_step"], d["clocks"])
PY
