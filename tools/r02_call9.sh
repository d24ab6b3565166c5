O=gpurun_out/r02i
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_all.txt 2>&1; echo "all tests exit $?: $(tail -1 $O/tests_all.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_all.txt | head -20
bench base
bench mt2 CIS_FORCE_MT128=2
bench kb32 CIS_HALO_STAGE_KB=32
bench kb64 CIS_HALO_STAGE_KB=64
bench sk32 CIS_SPLITK_NCTA=32
bench sk100 CIS_SPLITK_NCTA=100
ops base
head -14 $O/base.ops.txt
