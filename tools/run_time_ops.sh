CIS_PERSIST_MODE=0 timeout 300 python tools/time_ops.py > gpurun_out/time_ops_nopersist.txt 2>&1
timeout 300 python tools/time_ops.py > gpurun_out/time_ops_persist.txt 2>&1
tail -1 gpurun_out/time_ops_persist.txt
