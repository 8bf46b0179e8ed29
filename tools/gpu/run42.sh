timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_msgchn_step_gpu.py tests/test_msgchn_fullsize_gpu.py -q -m gpu -x 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 | tail -6
timeout 300 python bench.py --steps 200 --warmup 10 --no-extras 2>/dev/null | cut -c1-200
timeout 300 python bench.py --steps 200 --warmup 10 --no-extras --engine-opt tiled_up2_adj=0 2>/dev/null | cut -c1-200
