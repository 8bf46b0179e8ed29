PTTA_B200_LIB=tta_depth_completion_b200/lib/libptta_b200_stamps.so PTTA_ONE_STREAM=1 python tools/graph_stamps.py kitti > gpurun_out/r2_stamps_c.txt 2>&1
grep -A40 "start-to-start by kernel" gpurun_out/r2_stamps_c.txt | head -45
head -1 gpurun_out/r2_stamps_c.txt
