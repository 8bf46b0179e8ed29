python -m pytest tests/test_msgchn_step_gpu.py -m gpu -q -k "continual" 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-300
grep "continual" gpurun_out/parity_report.txt | grep -v weight | tail -8 | cut -c1-400
