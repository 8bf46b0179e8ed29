timeout 400 python bench.py --workload prepare_head --steps 50 --warmup 5 > gpurun_out/r2_bench_prepare_head.json 2> gpurun_out/r2_bench_prepare.err
cut -c1-1500 gpurun_out/r2_bench_prepare_head.json; tail -3 gpurun_out/r2_bench_prepare.err
timeout 400 python bench.py --workload prepare_init --steps 50 --warmup 5 > gpurun_out/r2_bench_prepare_init.json 2>> gpurun_out/r2_bench_prepare.err
cut -c1-1200 gpurun_out/r2_bench_prepare_init.json; tail -3 gpurun_out/r2_bench_prepare.err
