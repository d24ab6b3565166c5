O=gpurun_out/r02j
mkdir -p $O
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q --tb=short -p no:cacheprovider > $O/tests_nccl.txt 2>&1; echo "nccl test exit $?: $(tail -1 $O/tests_nccl.txt)"; grep -E "^FAILED|^ERROR|Error|assert" $O/tests_nccl.txt | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "2gpu: $(python tools/show_line.py $O/bench_2gpu.json)"; tail -3 $O/bench_2gpu.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload ensemble --no-cpu > $O/ens_2gpu.json 2> $O/ens_2gpu.err; echo "ensemble 2gpu: $(head -c 400 $O/ens_2gpu.json)"; tail -3 $O/ens_2gpu.err | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --workload ensemble > $O/ens_1gpu.json 2> $O/ens_1gpu.err; echo "ensemble 1gpu: $(head -c 600 $O/ens_1gpu.json)"; tail -3 $O/ens_1gpu.err | cut -c1-300
timeout 300 python bench.py --steps 50 --warmup 5 --workload gen_fwd > $O/gen_1gpu.json 2> $O/gen_1gpu.err; echo "gen_fwd 1gpu: $(head -c 600 $O/gen_1gpu.json)"; tail -3 $O/gen_1gpu.err | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "1gpu: $(python tools/show_line.py $O/bench_1gpu.json)"; tail -2 $O/bench_1gpu.err | cut -c1-300
(time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref_arm.json 2> $O/ref_arm.err); echo "ref arm: $(head -c 900 $O/ref_arm.json)"
