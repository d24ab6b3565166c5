CIS_SPLITK=0 timeout 300 python tools/time_ops.py > gpurun_out/time_ops_nosplit.txt 2>&1
timeout 300 python tools/time_ops.py > gpurun_out/time_ops_split.txt 2>&1
tail -3 gpurun_out/time_ops_split.txt
