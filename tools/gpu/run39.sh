nproc
timeout 600 python tools/pipeline_bench.py --steps 300 > gpurun_out/r2_pipeline.json 2> gpurun_out/r2_pipeline.err
cat gpurun_out/r2_pipeline.json; tail -5 gpurun_out/r2_pipeline.err
timeout 600 python tools/pipeline_bench.py --steps 300 --augment > gpurun_out/r2_pipeline_augment.json 2>> gpurun_out/r2_pipeline.err
cat gpurun_out/r2_pipeline_augment.json; tail -5 gpurun_out/r2_pipeline.err
