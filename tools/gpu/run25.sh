timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 > gpurun_out/r2_full_gpu_tests.log
tail -12 gpurun_out/r2_full_gpu_tests.log
timeout 300 python tools/prepare_timing.py > gpurun_out/r2_prepare_timing.json 2> gpurun_out/r2_prepare_timing.err
cat gpurun_out/r2_prepare_timing.json; tail -3 gpurun_out/r2_prepare_timing.err
