O=gpurun_out/r02o
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; tail -2 $O/$name.bench.err | cut -c1-200; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
timeout 120 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k "split_k_cluster" > $O/tests_cl.txt 2>&1; echo "cluster split tests exit $?: $(tail -1 $O/tests_cl.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_cl.txt | head
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_graph_parity_gpu.py -x -q --tb=short -p no:cacheprovider > $O/tests_conv.txt 2>&1; echo "conv tests exit $?: $(tail -1 $O/tests_conv.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_conv.txt | head
bench base
bench cl CIS_SPLITK_CLUSTER=1
bench cl_n100 CIS_SPLITK_CLUSTER=1 CIS_SPLITK_NCTA=100 CIS_SPLITK_MIN_UNITS=9
ops cl CIS_SPLITK_CLUSTER=1
head -6 $O/cl.ops.txt
