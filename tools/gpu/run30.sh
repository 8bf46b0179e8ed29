timeout 900 python -m pytest tests/test_nlspn_gpu.py tests/test_nlspn_net_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-300 | tail -8
timeout 120 python tools/prop_timing.py 2>&1 | tail -12
timeout 300 python bench.py --workload nlspn --steps 50 --warmup 5 --no-extras 2>/dev/null | cut -c1-260
