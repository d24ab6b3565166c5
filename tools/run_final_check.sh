timeout 900 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --steps 12 --warmup 4 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; python tools/show_bench.py gpurun_out/bench_check.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1]); print(d['clocks'], d['cpu_baseline'])"
