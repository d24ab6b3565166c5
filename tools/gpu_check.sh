# One GPU call that re-validates a change: GPU test suite, headline bench, per-op times.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_check.sh [tag] [ENV=VAL ...]'
TAG=${1:-check}; shift
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_all.txt 2>&1; echo "all tests exit $?: $(tail -1 $O/tests_all.txt)"; grep -E "^FAILED|^ERROR|Error" $O/tests_all.txt | head -20
env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "== bench: $(python tools/show_line.py $O/bench.json)"
env "$@" TIME_OPS_JSON=$O/ops.json timeout 400 python tools/time_ops.py > $O/ops.txt 2>&1; head -12 $O/ops.txt | grep -v checkpoint
