timeout 600 python -m pytest tests -q -m gpu --timeout 90 2>&1 | tail -2
timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -1 gpurun_out/bench_b.err | cut -c1-300; python tools/show_bench.py gpurun_out/bench_b.json
