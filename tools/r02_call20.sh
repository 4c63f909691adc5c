for cfg in "7 1" "7 2" "7 3" "7 4" "6 3" "14 2" "14 3"; do set -- $cfg
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-train --batch $1 --streams $2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $1 streams $2: ms', round(j['ms_per_step'],1), 'e2e', round(j['e2e']['ms_per_step'],1), 'clk', j['clocks']['sm_mhz'])"
done
