export CIS_MODES=G
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_halo_kernel.*32>" --launch-skip 13 --launch-count 1 -o gpurun_out/prof_halo32 -f python tools/ncu_step.py > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log
