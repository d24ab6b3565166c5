timeout 400 python tools/time_ops.py > gpurun_out/time_ops_graph.txt 2>&1
tail -1 gpurun_out/time_ops_graph.txt
