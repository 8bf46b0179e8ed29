# configs[2]: VOID 480x640 1layer, 8 independent shards, inner_iter 3 and 1
for it in 3 1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 8 --steps 100 --workload void --inner-iter $it 2>gpurun_out/r2_s8_void.err | tail -1 > gpurun_out/r2_s8_void_it$it.json
python -c "
import json; d=json.load(open('gpurun_out/r2_s8_void_it$it.json')); print('void8 inner', $it, round(d['value'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -5 gpurun_out/r2_s8_void.err
done
# configs[4]: shared model, batch 8 per GPU, 8 GPUs
for m in shared shared_nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29762 bench.py --gpus 8 --steps 50 --mode $m --batch 8 2>gpurun_out/r2_s8_${m}.err | tail -1 > gpurun_out/r2_s8_${m}.json
python -c "
import json; d=json.load(open('gpurun_out/r2_s8_${m}.json')); print('$m 8', round(d['value'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -5 gpurun_out/r2_s8_${m}.err
done
# configs[3]: NLSPN 8 shards
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29763 bench.py --gpus 8 --steps 30 --workload nlspn 2>gpurun_out/r2_s8_nlspn.err | tail -1 > gpurun_out/r2_s8_nlspn.json
python -c "
import json; d=json.load(open('gpurun_out/r2_s8_nlspn.json')); print('nlspn 8', round(d['value'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -5 gpurun_out/r2_s8_nlspn.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29764 tools/shared_check.py 2>&1 | grep "^{" > gpurun_out/r2_shared_check_8gpu.json; cat gpurun_out/r2_shared_check_8gpu.json | cut -c1-300
