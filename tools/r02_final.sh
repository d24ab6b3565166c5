# Final single-GPU evidence run of a round: test suite, the three bench workloads + the reference arm, per-op times, branch times, traces.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_final.sh'
O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/tests_all.txt 2>&1; echo "tests exit $?: $(tail -1 $O/tests_all.txt)"
timeout 600 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench: $(python tools/show_line.py $O/bench_1gpu.json)"
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 400 $O/bench_reference_arm.json
timeout 300 python bench.py --workload gen_fwd --steps 200 --warmup 20 > $O/bench_gen_fwd_1gpu.json 2> $O/bench_gen_fwd.err; tail -c 300 $O/bench_gen_fwd_1gpu.json
timeout 300 python bench.py --workload ensemble --steps 20 --warmup 5 > $O/bench_ensemble_1gpu.json 2> $O/bench_ensemble.err; tail -c 300 $O/bench_ensemble_1gpu.json
TIME_OPS_JSON=$O/ops.json timeout 400 python tools/time_ops.py > $O/ops.txt 2>&1; head -14 $O/ops.txt | grep -v checkpoint
timeout 300 python tools/time_branches.py 2>&1 | grep -v checkpoint > $O/branches.txt; cat $O/branches.txt
timeout 300 python tools/trace_conv.py 70 > $O/trace.txt 2>&1
timeout 300 python tools/trace_persist.py > $O/persist.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
