export CIS_MODES=G
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_halo_kernel.*128" --launch-skip 24 --launch-count 3 -o gpurun_out/prof_r01_halo128 -f python tools/ncu_step.py > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:"warp_costvol" --launch-skip 3 --launch-count 2 -o gpurun_out/prof_r01_costvol -f python tools/ncu_step.py > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
unset CIS_MODES
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_final.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; wc -l gpurun_out/launches_r01_final.csv
