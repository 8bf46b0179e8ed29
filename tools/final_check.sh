mkdir -p gpurun_out
set -o pipefail
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_full_gpu_tests_final.log
tail -3 gpurun_out/r2_full_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 200 --warmup 5 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/r2_bench_final.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
