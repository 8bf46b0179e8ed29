python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "tcgen05 or conv_tc" 2>&1 | tail -3
python tools/tc_timing.py 2>&1 | head -4
python bench.py --steps 100 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
