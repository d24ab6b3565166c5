# Round-2 GPU call 1: default suite + bench-size parity, every never-run kernel under its own timeout, hardware probes, the MMA-loop
# trace of the slowest launch shapes, per-op timings and A/B benches of the prepared switches.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_call1.sh'
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*"; }

stamp "default GPU suite (+ bench-size parity tests)"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_default.txt 2>&1
echo "default suite exit $?: $(tail -1 $O/tests_default.txt)"
grep -E "^FAILED|^ERROR" $O/tests_default.txt | head -20

bench() {   # name, env...
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu > $O/$name.bench.json 2> $O/$name.bench.err
  echo "== $name: $(python tools/show_line.py $O/$name.bench.json)"
}
ops() {     # name, env...
  name=$1; shift
  env "$@" TIME_OPS_JSON=$O/$name.ops.json timeout 400 python tools/time_ops.py > $O/$name.ops.txt 2>&1
  echo "== $name ops: $(grep -m1 'sum of warm' $O/$name.ops.txt)"
}
stamp "base bench + per-op times"
bench base
ops base
head -28 $O/base.ops.txt

stamp "MMA-loop trace"
CIS_LIB_NAME=libcis_b200_trace.so timeout 300 python tools/trace_conv.py 45 > $O/trace.txt 2>&1
echo "trace exit $?"; head -60 $O/trace.txt

stamp "A/B: planner-only switches"
bench deep148 CIS_DEEP_RING=148
bench deep296 CIS_DEEP_RING=296
bench thin8 CIS_HALO_SKIP_THIN=8
bench thin64 CIS_HALO_SKIP_THIN=64
bench wg2 CIS_WGRAD_CTAS_PER_SM=2
bench wg8 CIS_WGRAD_CTAS_PER_SM=8
bench persist0 CIS_PERSIST_MODE=0

stamp "experimental: functional API"
CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_functional_api_gpu.py -q --tb=short -p no:cacheprovider > $O/exp_functional.txt 2>&1
echo "functional exit $?: $(tail -1 $O/exp_functional.txt)"; grep -E "^FAILED|^ERROR" $O/exp_functional.txt | head

stamp "experimental: two-launch split-K"
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k two_launch > $O/exp_two_launch.txt 2>&1; then
  echo "two_launch OK: $(tail -1 $O/exp_two_launch.txt)"
  bench sk2_a CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=64 CIS_SPLITK_MIN_UNITS=18
  bench sk2_b CIS_SPLITK=2 CIS_SPLITK_MAX=8 CIS_SPLITK_NCTA=32 CIS_SPLITK_MIN_UNITS=18
  bench sk2_c CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=16 CIS_SPLITK_MIN_UNITS=27
  bench sk2_d CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=100 CIS_SPLITK_MIN_UNITS=36
  bench sk2_a_deep CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=64 CIS_SPLITK_MIN_UNITS=18 CIS_DEEP_RING=296
  ops sk2_a CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=64 CIS_SPLITK_MIN_UNITS=18
  python tools/ab_diff.py $O/base.ops.json $O/sk2_a.ops.json 3 | head -40
else
  echo "two_launch FAILED/timeout: $(tail -5 $O/exp_two_launch.txt)"
fi

stamp "experimental: weight-stationary persistent"
if CIS_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k weight_stationary > $O/exp_ws.txt 2>&1; then
  echo "ws OK: $(tail -1 $O/exp_ws.txt)"
  bench ws CIS_PERSIST_WS=1
else
  echo "ws FAILED/timeout: $(tail -5 $O/exp_ws.txt)"
fi

stamp "experimental: narrow n-tiles"
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k narrow_n_tiles > $O/exp_narrow.txt 2>&1; then
  echo "narrow OK: $(tail -1 $O/exp_narrow.txt)"
  bench bn32_5 CIS_SMALL_BN=32:5
  bench bn64_4 CIS_SMALL_BN=64:4
else
  echo "narrow FAILED/timeout: $(tail -5 $O/exp_narrow.txt)"
fi

stamp "hardware probes"
for p in umma_probe_mn umma_probe_noswz; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/$p tools/$p.cu > $O/$p.txt 2>&1 && timeout 60 /tmp/$p >> $O/$p.txt 2>&1
  echo "$p exit $?: $(grep -c -i ' ok' $O/$p.txt) ok lines, $(grep -c MISMATCH $O/$p.txt) mismatch lines"; tail -12 $O/$p.txt
done

stamp "experimental: halo wgrad"
if CIS_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k halo_wgrad > $O/exp_wgh.txt 2>&1; then
  echo "halo wgrad OK: $(tail -1 $O/exp_wgh.txt)"
  bench wgh CIS_WGRAD_HALO=1
  ops wgh CIS_WGRAD_HALO=1
  python tools/ab_diff.py $O/base.ops.json $O/wgh.ops.json 3 | head -30
else
  echo "halo wgrad FAILED/timeout: $(tail -8 $O/exp_wgh.txt)"
fi

stamp "experimental: 2-CTA cluster weight multicast"
if CIS_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_conv_engine_gpu.py -x -q --tb=short -p no:cacheprovider -k cluster_weight > $O/exp_cluster.txt 2>&1; then
  echo "cluster OK: $(tail -1 $O/exp_cluster.txt)"
  bench cluster CIS_HALO_CLUSTER=2
else
  echo "cluster FAILED/timeout: $(tail -5 $O/exp_cluster.txt)"
fi
stamp "done"
