timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 > gpurun_out/r2_full_gpu_tests.log
tail -8 gpurun_out/r2_full_gpu_tests.log
