timeout 600 python -m pytest tests/test_transforms_gpu.py tests/test_pipeline_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-500 | tail -10
