timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 > gpurun_out/r2_full_gpu_tests.log
tail -6 gpurun_out/r2_full_gpu_tests.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-extras 2>/dev/null | cut -c1-330
timeout 300 python bench.py --steps 200 --warmup 10 --no-extras --engine-opt fuse_bn_finalize=0 2>/dev/null | cut -c1-330
