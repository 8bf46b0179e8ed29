python -m pytest tests/test_ops_gpu.py tests/test_msgchn_fullsize_gpu.py tests/test_msgchn_step_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_e_tests.log
tail -4 gpurun_out/r2_e_tests.log
for o in "" "--engine-opt tc_stem_min_pixels=1000000000" "--engine-opt tc_stem_min_pixels=100000" "--engine-opt tc_stem_min_pixels=5000" "--engine-opt fuse_projpred=0"; do
python bench.py --steps 100 --no-extras $o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d.get('engine_options'), round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
done
