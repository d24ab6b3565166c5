O=gpurun_out/r02f
mkdir -p $O
bench() { name=$1; shift; env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err; echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"; tail -3 $O/$name.bench.err | cut -c1-200; }
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_all.txt 2>&1; echo "all tests exit $?: $(tail -1 $O/tests_all.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_all.txt | head -20
bench base
bench nopipe CIS_PIPELINE=0
bench wg2 CIS_WGRAD_CTAS_PER_SM=2
bench wg3 CIS_WGRAD_CTAS_PER_SM=3
bench nowgh CIS_WGRAD_HALO=0
