O=gpurun_out/r02e
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_functional_api_gpu.py tests/test_graph_parity_gpu.py -q --tb=short -p no:cacheprovider > $O/tests_conv.txt 2>&1; echo "conv tests exit $?: $(tail -1 $O/tests_conv.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_conv.txt | head -20
CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -q --tb=short -p no:cacheprovider -k halo_wgrad > $O/exp_wgh.txt 2>&1; echo "halo wgrad exit $?: $(tail -1 $O/exp_wgh.txt)"; grep -E "^FAILED|^ERROR|Error" $O/exp_wgh.txt | head
bench base
bench persist0 CIS_PERSIST_MODE=0
bench persist2 CIS_PERSIST_MODE=2
bench pmin148 CIS_PERSIST_MIN_TILES=148
bench pmin600 CIS_PERSIST_MIN_TILES=600
bench wgh CIS_WGRAD_HALO=1
ops base
head -12 $O/base.ops.txt
