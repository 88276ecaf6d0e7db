#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for the bench command (run under gpurun).
# usage: tools/profile_cmd.sh <tag> [kernel regex] [extra bench args]
TAG=$1; shift
KRE=${1:-estep_evo_bsc|bsc_mstep}; shift
CMD="python bench.py --n 131072 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other $@"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s ${SKIP:-2} -c ${COUNT:-2} -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "profile rc=$?"
tail -2 gpurun_out/plain_$TAG.log | cut -c1-600
