timeout 600 python -m pytest tests/test_transforms_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 | tail -8
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_sanitizer_memcheck_step.log python -m pytest tests/test_msgchn_step_gpu.py -q -m gpu -k "fused_bn_finalize or graph_replay" -x 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-300 | tail -4
tail -3 gpurun_out/r2_sanitizer_memcheck_step.log
