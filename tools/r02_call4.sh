O=gpurun_out/r02d
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; }
ops() { name=$1; shift; env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1; echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"; }
CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k "two_launch or narrow_n_tiles" > $O/exp_sk.txt 2>&1; echo "two_launch+narrow exit $?: $(tail -1 $O/exp_sk.txt)"; grep -E "^FAILED|^ERROR|Error" $O/exp_sk.txt | head
CIS_TEST_EXPERIMENTAL=1 CIS_HALO_MIN_UTIL=0.15 timeout 300 python -m pytest tests/test_conv_engine_gpu.py tests/test_graph_parity_gpu.py -x -q --tb=short -p no:cacheprovider > $O/tests_util.txt 2>&1; echo "util .15 tests exit $?: $(tail -1 $O/tests_util.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_util.txt | head
bench base
bench util2 CIS_HALO_MIN_UTIL=0.2
bench util2_sk_a CIS_HALO_MIN_UTIL=0.2 CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=64 CIS_SPLITK_MIN_UNITS=18
bench util2_sk_b CIS_HALO_MIN_UTIL=0.2 CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=32 CIS_SPLITK_MIN_UNITS=18
bench util2_sk_c CIS_HALO_MIN_UTIL=0.2 CIS_SPLITK=2 CIS_SPLITK_MAX=8 CIS_SPLITK_NCTA=100 CIS_SPLITK_MIN_UNITS=27
bench util2_sk_d CIS_HALO_MIN_UTIL=0.2 CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=16 CIS_SPLITK_MIN_UNITS=9
bench util4 CIS_HALO_MIN_UTIL=0.4
ops util2_sk_a CIS_HALO_MIN_UTIL=0.2 CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=64 CIS_SPLITK_MIN_UNITS=18
head -12 $O/util2_sk_a.ops.txt
