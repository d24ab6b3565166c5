# Round-2 ncu evidence (one GPU, never multi-rank): launch list + tensor-pipe % of one G and one R step, --set full captures of the
# dominant launches (indices from the launch list), cost volume at batch 64.  Results are summarised into profiles/ by hand
# (ncu -i ... --page raw --csv) -- gpurun_out/ is scratch.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_profile.sh'
O=gpurun_out/r02p
mkdir -p $O
export CIS_PIPELINE=0
CIS_OPS_JSON=$O/ops.json timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/launches.csv python tools/ncu_step.py > $O/ncu_step.log 2>&1; echo "launch list exit $?: $(wc -l < $O/launches.csv) lines"
python tools/ncu_table.py $O/launches.csv $O/ops.json > $O/table.txt 2>&1; head -64 $O/table.txt
# index of the dominant conv_halo_kernel<128> launch (grid 480 = PWC-Net level-2 context conv 576->128) among its kind, first step
SKIP=$(python - <<'PY'
import csv
lines = [l for l in open('gpurun_out/r02p/launches.csv') if l.startswith('"')]
per = {}
for r in csv.DictReader(lines):
    d = per.setdefault(r['ID'], dict(name=r['Kernel Name'], grid=r['Grid Size']))
    if r['Metric Name'] == 'gpu__time_duration.sum':
        d['t'] = float(r['Metric Value'].replace(',', ''))
h = [d for d in per.values() if 'conv_halo_kernel<128>' in d['name']]
best = max(range(min(len(h), 80)), key=lambda i: h[i].get('t', 0))
print(best)
PY
)
echo "dominant halo128 launch index: $SKIP"
export CIS_MODES=G
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_halo_kernel.*128" --launch-skip $SKIP --launch-count 1 -o $O/prof_halo128_dominant -f python tools/ncu_step.py > $O/ncu_full1.log 2>&1; tail -1 $O/ncu_full1.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_halo_persist_kernel" --launch-skip 10 --launch-count 12 -o $O/prof_persist -f python tools/ncu_step.py > $O/ncu_full2.log 2>&1; tail -1 $O/ncu_full2.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:"conv_wgrad" --launch-skip 2 --launch-count 4 -o $O/prof_wgrad -f python tools/ncu_step.py > $O/ncu_full3.log 2>&1; tail -1 $O/ncu_full3.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:"conv_igemm_kernel" --launch-skip 4 --launch-count 3 -o $O/prof_gather -f python tools/ncu_step.py > $O/ncu_full4.log 2>&1; tail -1 $O/ncu_full4.log
unset CIS_MODES
timeout 120 python tools/costvol_roofline.py > $O/costvol_b64.json 2> $O/costvol_b64.err; cat $O/costvol_b64.json
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"warp_costvol" --launch-skip 4 --launch-count 2 -o $O/prof_costvol_b64 -f python tools/costvol_roofline.py > $O/ncu_full5.log 2>&1; tail -1 $O/ncu_full5.log
