for s in 1 2 3; do
  timeout 300 python bench.py --steps 3 --warmup 3 --streams $s --no-cpu-baseline --no-train 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('streams $s', 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'], 'frac_sust', round(d['roofline']['whole_step_frac_of_sustained'],3))
"
done
