O=gpurun_out/r02m
mkdir -p $O
CIS_LIB_NAME=libcis_b200_trace.so timeout 300 python tools/trace_persist.py > $O/trace_persist.txt 2>&1; grep -v checkpoint $O/trace_persist.txt | head -120
