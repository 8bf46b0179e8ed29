timeout 600 python -m pytest tests/test_transforms_gpu.py -q -m gpu 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-400 > gpurun_out/r2_transforms_tests.log
tail -10 gpurun_out/r2_transforms_tests.log
timeout 300 python tools/prepare_timing.py > gpurun_out/r2_prepare_timing.json 2> gpurun_out/r2_prepare_timing.err
cat gpurun_out/r2_prepare_timing.json; tail -3 gpurun_out/r2_prepare_timing.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_augment_kernels.csv -k regex:"photo_|flip_kernel" -c 40 python tools/prepare_timing.py --batch 4 --iters 2 > gpurun_out/r2_augment_ncu.log 2>&1
tail -2 gpurun_out/r2_augment_ncu.log
