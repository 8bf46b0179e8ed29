python tools/small_kernels_timing.py 2>&1 | tee gpurun_out/r2_small_timing.txt
ncu --set full --clock-control none --import-source on -k regex:"stem_tc|tc_head" -c 4 -o gpurun_out/r2_stemtc python tools/small_kernels_timing.py > gpurun_out/r2_stemtc_ncu.log 2>&1
tail -3 gpurun_out/r2_stemtc_ncu.log
