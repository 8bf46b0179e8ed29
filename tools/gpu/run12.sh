python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 100 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
