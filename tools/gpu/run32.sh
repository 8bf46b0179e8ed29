timeout 900 python -m pytest tests/test_prepare_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 | tail -10
timeout 300 python tools/prepare_timing.py --batch 1 > gpurun_out/r2_prepare_timing_bm256.json 2> gpurun_out/r2_prepare_timing.err
cat gpurun_out/r2_prepare_timing_bm256.json; tail -3 gpurun_out/r2_prepare_timing.err
