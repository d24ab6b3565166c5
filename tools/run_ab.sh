# One GPU call that measures the round-2 experiment switches of DESIGN.md section 6 against the default build:
# per-launch CUDA-graph-replay times (tools/time_ops.py) + the whole-step bench for each configuration.
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/run_ab.sh'      (about 14 configurations x ~100 s + the gated parity tests)
# Experimental kernels (two-launch split-K, weight-stationary persistent conv, halo wgrad) are parity-tested first and skipped on failure.
mkdir -p gpurun_out/ab
run() {   # name, env assignments...
  name=$1; shift
  env "$@" TIME_OPS_JSON=gpurun_out/ab/$name.json timeout 200 python tools/time_ops.py > gpurun_out/ab/$name.txt 2>&1
  env "$@" timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ab/$name.bench.json 2> gpurun_out/ab/$name.bench.err
  echo "== $name: $(head -c 300 gpurun_out/ab/$name.bench.json | python -c 'import sys,json; d=json.loads(sys.stdin.readline() or "{}"); print(d.get("value"), d.get("ms_per_step"))' 2>/dev/null)  $(grep -m1 "sum of warm" gpurun_out/ab/$name.txt)"
}
run base
run e1_thin8 CIS_HALO_SKIP_THIN=8
run e1_thin64 CIS_HALO_SKIP_THIN=64
# two-launch split-K is experimental: validate it first, skip its timing runs if the parity test fails
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q -k two_launch > gpurun_out/ab/two_launch_test.txt 2>&1; then
  run e2_sk2_16 CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=8 CIS_SPLITK_MIN_UNITS=32      # only the 6x10 / 4x7 maps (python tools/plan_report.py previews the selection)
  run e2_sk2_8 CIS_SPLITK=2 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=32 CIS_SPLITK_MIN_UNITS=32       # + 12x20 and 24x40 halo layers
  run e2_sk2_wide CIS_SPLITK=2 CIS_SPLITK_MAX=8 CIS_SPLITK_NCTA=96 CIS_SPLITK_MIN_UNITS=45     # + the long 48x80 / 32x56 layers
else
  echo "two-launch split-K parity FAILED:"; tail -5 gpurun_out/ab/two_launch_test.txt
fi
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q -k weight_stationary > gpurun_out/ab/ws_test.txt 2>&1; then
  run e1_ws CIS_PERSIST_WS=1
else
  echo "weight-stationary persistent kernel parity FAILED:"; tail -5 gpurun_out/ab/ws_test.txt
fi
# halo-resident wgrad: hardware probe first, then parity, then timing
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe_mn tools/umma_probe_mn.cu > gpurun_out/ab/probe_mn.txt 2>&1 && timeout 60 /tmp/umma_probe_mn >> gpurun_out/ab/probe_mn.txt 2>&1
echo "umma_probe_mn: $(grep -c ' ok ' gpurun_out/ab/probe_mn.txt) ok, $(grep -c MISMATCH gpurun_out/ab/probe_mn.txt) mismatch"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe_noswz tools/umma_probe_noswz.cu > gpurun_out/ab/probe_noswz.txt 2>&1 && timeout 60 /tmp/umma_probe_noswz >> gpurun_out/ab/probe_noswz.txt 2>&1
echo "umma_probe_noswz: $(grep -c ': ok' gpurun_out/ab/probe_noswz.txt) ok, $(grep -c MISMATCH gpurun_out/ab/probe_noswz.txt) mismatch"
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q -k halo_wgrad > gpurun_out/ab/wgh_test.txt 2>&1; then
  run e4_wgh CIS_WGRAD_HALO=1
else
  echo "halo wgrad parity FAILED:"; tail -5 gpurun_out/ab/wgh_test.txt
fi
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q -k cluster_weight > gpurun_out/ab/cl_test.txt 2>&1; then
  run e3_cluster CIS_HALO_CLUSTER=2
else
  echo "cluster weight multicast parity FAILED:"; tail -5 gpurun_out/ab/cl_test.txt
fi
if CIS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_engine_gpu.py -x -q -k narrow_n_tiles > gpurun_out/ab/narrow_test.txt 2>&1; then
  run e2_bn32 CIS_SMALL_BN=32:5
  run e2_bn32_l4 CIS_SMALL_BN=32:4
  run e2_bn64_l4 CIS_SMALL_BN=64:4
else
  echo "narrow n-tile parity FAILED:"; tail -5 gpurun_out/ab/narrow_test.txt
fi
run e2_sk1_16 CIS_SPLITK=1 CIS_SPLITK_MAX=16 CIS_SPLITK_NCTA=8 CIS_SPLITK_MIN_UNITS=32
for v in e1_thin8 e1_thin64 e1_ws e3_cluster e4_wgh e2_bn32 e2_bn32_l4 e2_bn64_l4 e2_sk2_16 e2_sk2_8 e2_sk2_wide e2_sk1_16; do
  [ -f gpurun_out/ab/$v.json ] || continue
  echo "---- $v vs base"; python tools/ab_diff.py gpurun_out/ab/base.json gpurun_out/ab/$v.json 3 | head -25
done
