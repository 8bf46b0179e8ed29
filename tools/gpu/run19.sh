timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29771 tools/shared_check.py > gpurun_out/r2_shared_check_8gpu.log 2>&1
grep "^{" gpurun_out/r2_shared_check_8gpu.log | cut -c1-900
grep -i "assert\|Error" gpurun_out/r2_shared_check_8gpu.log | head -8 | cut -c1-300
