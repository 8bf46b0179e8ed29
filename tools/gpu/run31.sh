timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -12
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
cut -c1-600 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
