#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench, then the ncu launch list and one --set full capture
# of the same command.  usage: tools/gpu_profile.sh <tag> [bench args...]
set -u
tag=$1; shift
args="--steps 2 --warmup 3 $*"
mkdir -p gpurun_out
python bench.py $args > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || { echo "bench failed"; tail -20 gpurun_out/bench_${tag}.err; exit 1; }
cat gpurun_out/bench_${tag}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py $args --no-cpu-baseline > gpurun_out/ncu_list_${tag}.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1_fq_pass_a|k1_scan|k1_fq_pass_c|k2_hash' -s 132 -c 8 \
    -o gpurun_out/prof_${tag} -f python bench.py $args --no-cpu-baseline --no-dist > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'rs_|k3_' -s 60 -c 12 \
    -o gpurun_out/prof_sel_${tag} -f python bench.py $args --no-cpu-baseline --no-dist > gpurun_out/ncu_full_sel_${tag}.log 2>&1
echo "ncu full sel rc=$?"
ls -la gpurun_out
