timeout 600 python -m pytest tests/test_prepare_gpu.py -q -m gpu -k "checkpoint_round_trip" 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29791 bench.py --gpus 2 --steps 100 --warmup 5 2>/dev/null | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29792 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference 2>/dev/null | cut -c1-300
