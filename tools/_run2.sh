python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_hl.json | cut -c1-1500
python tools/time_kernels.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_count_lean" -s 8 -c 6 -o gpurun_out/prof_hl3 python tools/time_kernels.py > gpurun_out/ncu_hl3.log 2>&1
