set -x
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1f.json | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_count_lean" -s 2 -c 2 -o gpurun_out/prof_r1f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
python tools/time_ops.py > gpurun_out/ops_r1f.log 2>&1
python tools/time_api.py 2000000 > gpurun_out/api_r1f.log 2>&1
