python -m pytest tests/test_nlspn_net_gpu.py tests/test_nlspn_gpu.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_f_tests.log
tail -12 gpurun_out/r2_f_tests.log
python -c "
import __graft_entry__ as g
g.smoke()
" 2>&1 | tail -6
python bench.py --workload nlspn --steps 30 2>/dev/null > gpurun_out/r2_f_bench_nlspn.json; python -c "
import json; d=json.load(open('gpurun_out/r2_f_bench_nlspn.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'], d['clocks'])"
