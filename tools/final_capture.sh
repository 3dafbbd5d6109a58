# Round-end evidence on one B200 (run through gpurun): the bench line, the ncu launch list, one full capture of the
# two query-side kernels, per-op and public-API timings.  Every ncu command runs only after the same command has
# exited 0 without ncu; numbers printed under ncu are never bench values.
set -x
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1f.json | cut -c1-300
A="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$A > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv $A > gpurun_out/ncu1.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_count_lean" -s 2 -c 2 -o gpurun_out/prof_r1f $B > gpurun_out/ncu2.log 2>&1
python tools/time_ops.py > gpurun_out/ops_r1f.log 2>&1
python tools/time_api.py 2000000 > gpurun_out/api_r1f.log 2>&1
