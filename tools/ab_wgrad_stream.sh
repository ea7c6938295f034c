for rep in 1 2; do for f in 0 1; do
CRIS_B200_WGRAD_STREAM=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-incumbent 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WGRAD_STREAM=$f', round(d['value'],1), round(d['ms_per_step'],2), d['clocks'])"
done; done
