timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29781 tools/shared_check.py --stage init > gpurun_out/r2_shared_check_init_2gpu.log 2>&1
grep "^{" gpurun_out/r2_shared_check_init_2gpu.log | cut -c1-900; grep -i "assert\|Error" gpurun_out/r2_shared_check_init_2gpu.log | head -6 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29782 tools/shared_check.py --stage head > gpurun_out/r2_shared_check_head_2gpu.log 2>&1
grep "^{" gpurun_out/r2_shared_check_head_2gpu.log | cut -c1-900; grep -i "assert\|Error" gpurun_out/r2_shared_check_head_2gpu.log | head -6 | cut -c1-300
