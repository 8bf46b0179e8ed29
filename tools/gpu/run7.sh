timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 tools/shared_check.py > gpurun_out/r2_shared_check.log 2>&1
grep -v "^W\|^\*\*\*\|^$" gpurun_out/r2_shared_check.log | grep -i "assert\|Error\|error\|{" | head -20
