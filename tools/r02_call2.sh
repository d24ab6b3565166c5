O=gpurun_out/r02b
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; tail -2 $O/$name.bench.err | cut -c1-300; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
timeout 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_kernels_gpu.py -x -q --tb=short -p no:cacheprovider > $O/tests_conv.txt 2>&1; echo "conv tests exit $?: $(tail -1 $O/tests_conv.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_conv.txt | head
bench base
ops base
head -8 $O/base.ops.txt
python tools/ab_diff.py gpurun_out/r02/base.ops.json $O/base.ops.json 5 | head -60
bench g1 CIS_HALO_G=1
bench kb16 CIS_HALO_STAGE_KB=16
bench kb48 CIS_HALO_STAGE_KB=48
bench kb64 CIS_HALO_STAGE_KB=64
bench persist1 CIS_PERSIST_MODE=1
CIS_LIB_NAME=libcis_b200_trace.so timeout 300 python tools/trace_conv.py 30 > $O/trace.txt 2>&1; head -70 $O/trace.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_all.txt 2>&1; echo "all tests exit $?: $(tail -1 $O/tests_all.txt)"; grep -E "^FAILED|^ERROR" $O/tests_all.txt | head -20
