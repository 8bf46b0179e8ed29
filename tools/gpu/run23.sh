timeout 300 python tools/prepare_timing.py > gpurun_out/r2_prepare_timing.json 2> gpurun_out/r2_prepare_timing.err
cat gpurun_out/r2_prepare_timing.json; tail -3 gpurun_out/r2_prepare_timing.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_prepare_kernels.csv -k regex:"gemm_tn|photo_|flip_kernel|rotate_kernel|resize_crop|l2_loss|cos_loss|ema_update" -c 60 python tools/prepare_timing.py --batch 1 --iters 2 > gpurun_out/r2_prepare_ncu.log 2>&1
tail -3 gpurun_out/r2_prepare_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc_kernel -c 3 -o gpurun_out/r2_gemm_tn python tools/prepare_timing.py --batch 1 --iters 2 > gpurun_out/r2_gemm_tn_ncu.log 2>&1
tail -2 gpurun_out/r2_gemm_tn_ncu.log
