timeout 600 python -m pytest tests/test_transforms_gpu.py -q -m gpu -s 2>&1 | grep -E "^E  |FAILED|passed|failed|fraction" | cut -c1-400 > gpurun_out/r2_transforms_tests.log
tail -30 gpurun_out/r2_transforms_tests.log
