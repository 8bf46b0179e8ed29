python -m pytest tests/test_shared_gpu.py -m gpu -q -x 2>&1 | tail -15
for m in shared shared_nccl; do

timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 50 --mode $m --batch 8 2>gpurun_out/r2_g_${m}.err | tail -1 > gpurun_out/r2_g_bench_${m}_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_g_bench_${m}_2gpu.json')); print('$m', round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'], d['cuda_graph'])" || tail -5 gpurun_out/r2_g_${m}.err
done
