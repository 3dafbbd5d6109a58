# Round-2 evidence on one B200 (run through gpurun): the bench line, the ncu launch lists (join step; C4 / C5 operations),
# full captures of the dominant kernels, public-API timings.  Every ncu command runs only after the same command has
# exited 0 without ncu; numbers printed under ncu are never bench values.  Usage: bash tools/final_capture_r2.sh <tag>
T=${1:-r2w}
set -x
python bench.py --steps 10 --warmup 3 2>gpurun_out/${T}_bench.err > gpurun_out/${T}_bench.json
A="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --ops count,overlap"
$A > gpurun_out/${T}_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_join.csv $A > gpurun_out/${T}_ncu1.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --ops cluster,merge,rle,nearest --ops-steps 2"
$B > gpurun_out/${T}_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_ops.csv $B > gpurun_out/${T}_ncu2.log 2>&1
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ops none"
$C > gpurun_out/${T}_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_join_pipe" -s 3 -c 1 -o gpurun_out/${T}_join $C > gpurun_out/${T}_ncu3.log 2>&1
D="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --ops cluster --ops-steps 2"
ncu --set full --clock-control none --import-source on -k regex:"k_rs_sweep|k_hy_buckets" -s 4 -c 4 -o gpurun_out/${T}_unary $D > gpurun_out/${T}_ncu4.log 2>&1
E2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --ops rle,overlap --ops-steps 2"
ncu --set full --clock-control none --import-source on -k regex:"k_rd_dense|k_rd_pass_a|k_join_any" -s 3 -c 4 -o gpurun_out/${T}_rle $E2 > gpurun_out/${T}_ncu5.log 2>&1
python tools/time_api.py 2000000 > gpurun_out/${T}_api.log 2>&1
