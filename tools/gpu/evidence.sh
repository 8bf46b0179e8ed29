# round-2 evidence: launch lists with DRAM bytes (eager step, serialised under ncu), full capture of the roofline kernel, timing tables, bench lines
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_step_b1.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras > gpurun_out/r2_ncu_b1.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_step_b8.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras --batch 8 > gpurun_out/r2_ncu_b8.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_step_nlspn.csv python bench.py --workload nlspn --steps 1 --warmup 3 --no-graph --no-extras > gpurun_out/r2_ncu_nlspn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -c 8 -o gpurun_out/r2_tc python tools/tc_profile_target.py > gpurun_out/r2_tc_ncu.log 2>&1
python tools/small_kernels_timing.py > gpurun_out/r2_small_timing_final.txt 2>&1
python tools/gemm_timing.py > gpurun_out/r2_gemm_timing.txt 2>&1
PTTA_B200_LIB=tta_depth_completion_b200/lib/libptta_b200_stamps.so PTTA_ONE_STREAM=1 python tools/graph_stamps.py kitti > gpurun_out/r2_stamps_final_1stream.txt 2>&1
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --workload void > gpurun_out/r2_bench_void.json 2>> gpurun_out/r2_bench.err
python bench.py --workload nlspn --steps 50 > gpurun_out/r2_bench_nlspn.json 2>> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 5 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
python bench.py --batch 8 --steps 50 --no-extras > gpurun_out/r2_bench_b8.json 2>> gpurun_out/r2_bench.err
tail -c 600 gpurun_out/r2_bench.json
