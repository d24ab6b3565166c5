O=gpurun_out/r02q
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; tail -2 $O/$name.bench.err | cut -c1-200; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_all.txt 2>&1; echo "all tests exit $?: $(tail -1 $O/tests_all.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_all.txt | head -20
bench base
bench nogroup CIS_GROUP_PARITY=0
bench prio CIS_PIPE_PRIO=1
bench base2
ops base
head -8 $O/base.ops.txt
