#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench, then the ncu launch list and --set full captures
# of the same command.  usage: tools/gpu_profile.sh <tag> [extra bench args...]
set -u
tag=$1; shift
mkdir -p gpurun_out
full="--steps 3 --warmup 3 $*"
prof="--steps 2 --warmup 3 --no-cpu-baseline --no-dist $*"
python bench.py $full > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || { echo "bench failed"; tail -20 gpurun_out/bench_${tag}.err; exit 1; }
cat gpurun_out/bench_${tag}.json
python bench.py $prof > gpurun_out/bench_prof_${tag}.json 2> gpurun_out/bench_prof_${tag}.err || { echo "bench (profile args) failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py $prof > gpurun_out/ncu_list_${tag}.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1q_|k1_fa|k2_hash' -s 40 -c 6 \
    -o gpurun_out/prof_${tag} -f python bench.py $prof > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'bs_|k3t|k4_|rs_' -s 10 -c 16 \
    -o gpurun_out/prof_sel_${tag} -f python bench.py $prof > gpurun_out/ncu_full_sel_${tag}.log 2>&1
echo "ncu full sel rc=$?"
ls -la gpurun_out
