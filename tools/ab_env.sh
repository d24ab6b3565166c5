# A/B of environment switches on the headline bench: bash tools/ab_env.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...   (one bench run per quoted set)
TAG=$1; shift
mkdir -p gpurun_out/$TAG
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/$TAG/bench_$i.json 2> gpurun_out/$TAG/bench_$i.err
  echo "[$cfg] $(python tools/show_line.py gpurun_out/$TAG/bench_$i.json)"
done
