timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest tests/test_transforms_gpu.py tests/test_prepare_gpu.py -q -m gpu -k "fixture or scalar_and_vector or gemm_tn or fused_preparation or graph_replay" -x 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-300 | tail -6
echo "sanitizer rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r2_sanitizer_racecheck.log python -m pytest tests/test_transforms_gpu.py -q -m gpu -k "scalar_and_vector" 2>&1 | grep -E "FAILED|ERROR|passed|failed" | tail -3
tail -4 gpurun_out/r2_sanitizer_racecheck.log
