python -m pytest tests/test_msgchn_step_gpu.py tests/test_msgchn_fullsize_gpu.py tests/test_checkpoint_gpu.py tests/test_shared_gpu.py -m gpu -q 2>&1 | tail -6
for o in "" "--engine-opt fuse_enc_sums=0"; do
python bench.py --steps 100 --no-extras $o 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d.get('engine_options'), round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])"
done
PTTA_B200_LIB=tta_depth_completion_b200/lib/libptta_b200_stamps.so PTTA_ONE_STREAM=1 python tools/graph_stamps.py kitti > gpurun_out/r2_stamps_final_1stream.txt 2>&1
grep -A24 "start-to-start by kernel" gpurun_out/r2_stamps_final_1stream.txt; head -1 gpurun_out/r2_stamps_final_1stream.txt
