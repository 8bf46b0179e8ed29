python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -5
python -m pytest tests/test_msgchn_step_gpu.py tests/test_msgchn_fullsize_gpu.py tests/test_nlspn_net_gpu.py -m gpu -q -x 2>&1 | tail -4
python tools/gemm_timing.py 2>&1 | tail -8
python bench.py --steps 100 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
