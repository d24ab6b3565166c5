O=gpurun_out/r02h
mkdir -p $O
timeout 300 python tools/time_branches.py 2>&1 | grep -v "checkpoint" | tee $O/branches.txt
