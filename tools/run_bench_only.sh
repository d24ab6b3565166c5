timeout 300 python -m pytest tests/test_conv_engine_gpu.py tests/test_graph_parity_gpu.py -q -x --timeout 60 2>&1 | tail -2
timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -1 gpurun_out/bench_b.err | cut -c1-300; python tools/show_bench.py gpurun_out/bench_b.json
