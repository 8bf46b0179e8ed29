python -m pytest tests/test_ops_gpu.py tests/test_msgchn_fullsize_gpu.py tests/test_msgchn_step_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_d_tests.log
tail -4 gpurun_out/r2_d_tests.log
for o in "" "--engine-opt tc_head_min_pixels=1000000000" "--engine-opt tc_min_pixels=1500 --engine-opt tc_s2_min_pixels=1500 --engine-opt tc_t2_min_pixels=1000" "--engine-opt tc_min_pixels=6000 --engine-opt tc_s2_min_pixels=6000 --engine-opt tc_t2_min_pixels=1500" "--engine-opt tc_head_min_pixels=100000"; do
python bench.py --steps 100 --no-extras $o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d.get('engine_options'), round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
done
