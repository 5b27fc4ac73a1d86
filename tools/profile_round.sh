#!/bin/bash
# tools/profile_round.sh <tag>   -- the profiling pass of a round, run on the GPU box (gpurun).  Writes into gpurun_out/:
#   <tag>_launches.csv        ncu launch list (gpu__time_duration.sum) of the default bench command
#   <tag>_<workload>.ncu-rep  one --set full capture of the step kernel per workload (after warm-up / spin-up)
# Summaries for profiles/ are made afterwards where the reports were copied to (tools/profile_summaries.sh).
tag=${1:-rNN}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -c 1 -f"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/${tag}_launch_bench.log 2>&1
echo "launch list rc=$?"
for wl in dambreak4096 dambreak4096-f32 dambreak4096-mh dambreak4096-mh-f32 dambreak4096-inertial dambreak4096-inertial-f32; do
    timeout 300 $NCU -k regex:step_ -o gpurun_out/${tag}_${wl} python tools/run_short.py $wl 2 0 4096 0.3 > gpurun_out/${tag}_ncu_${wl}.log 2>&1
    echo "$wl rc=$?"
done
timeout 600 $NCU -k regex:step_ -o gpurun_out/${tag}_pluvial16384 python tools/run_short.py pluvial16384 2 0 16384 2.2 > gpurun_out/${tag}_ncu_pluvial16384.log 2>&1
echo "pluvial16384 rc=$?"
timeout 600 $NCU -k regex:step_ -o gpurun_out/${tag}_river32768 python tools/run_short.py river32768 2 0 0 0.5 > gpurun_out/${tag}_ncu_river32768.log 2>&1
echo "river32768 rc=$?"
