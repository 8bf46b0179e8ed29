python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "stem" 2>&1 | tail -4
python tools/small_kernels_timing.py 2>&1 | head -4 | tee gpurun_out/r2_small_timing2.txt
for o in "" "--engine-opt tc_stem_min_pixels=1000000000"; do
python bench.py --steps 100 --no-extras $o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d.get('engine_options'), round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
done
