timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in C4 C2; do timeout 200 python bench.py --workload $w --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$w', 'ms %.4f' % d['ms_per_step'], 'Mrays/s %.0f' % d['value'])"; done
