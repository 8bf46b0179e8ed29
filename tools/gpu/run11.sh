python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "upsampled or tcgen05" 2>&1 | tail -5
python -m pytest tests/test_msgchn_step_gpu.py tests/test_msgchn_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -4
for o in "" "--engine-opt fuse_up2=0"; do
python bench.py --steps 100 --no-extras $o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d.get('engine_options'), round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
done
