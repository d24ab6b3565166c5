O=gpurun_out/r02n
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_graph_parity_gpu.py -x -q --tb=short -p no:cacheprovider > $O/tests_conv.txt 2>&1; echo "conv tests exit $?: $(tail -1 $O/tests_conv.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_conv.txt | head
bench base
ops base
head -6 $O/base.ops.txt
CIS_LIB_NAME=libcis_b200_trace.so timeout 300 python tools/trace_persist.py > $O/trace_persist.txt 2>&1; grep -v checkpoint $O/trace_persist.txt | head -90
