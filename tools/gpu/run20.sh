rm -f gpurun_out/prepare_parity_report.txt
timeout 900 python -m pytest tests/test_prepare_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|passed|failed|^prep_|^fullsize|^head training" | cut -c1-420 > gpurun_out/r2_prepare_tests.log
cat gpurun_out/r2_prepare_tests.log | tail -60
