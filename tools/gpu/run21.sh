rm -f gpurun_out/prepare_parity_report.txt
timeout 900 python -m pytest tests/test_prepare_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-300 > gpurun_out/r2_prepare_tests.log
tail -12 gpurun_out/r2_prepare_tests.log
timeout 300 python tools/prepare_timing.py > gpurun_out/r2_prepare_timing.json 2> gpurun_out/r2_prepare_timing.err
cat gpurun_out/r2_prepare_timing.json; tail -5 gpurun_out/r2_prepare_timing.err
