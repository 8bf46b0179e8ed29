timeout 600 python -m pytest tests/test_transforms_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|ERROR|passed|failed" | cut -c1-400 | tail -8
timeout 600 python tools/pipeline_bench.py --steps 300 --augment > gpurun_out/r2_pipeline_augment.json 2> gpurun_out/r2_pipeline.err
cat gpurun_out/r2_pipeline_augment.json; tail -5 gpurun_out/r2_pipeline.err
