#!/bin/bash
# one ncu --set full capture of the tower kernel (after the same command exited 0 without ncu).  Usage: gpu_ncu_tower.sh <tag> [bench args]
tag=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-sims 2 $@"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:tower_tcgen05 -s 6 -c 1 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -n 3 gpurun_out/ncu_full_$tag.log
