timeout 600 python -m pytest tests -q -m gpu --timeout 90 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.err | cut -c1-200; cat gpurun_out/bench_ref.json | cut -c1-700
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -1 gpurun_out/bench_r01_final.err | cut -c1-300; cat gpurun_out/bench_r01_final.json | cut -c1-3000
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
