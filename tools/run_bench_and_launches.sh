set -x
python bench.py --steps 8 --warmup 4 --no-cpu > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_r01_a.err; tail -3 gpurun_out/bench_r01_a.err; cat gpurun_out/bench_r01_a.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/launches_r01.csv
