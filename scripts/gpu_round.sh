#!/bin/bash
# One GPU session: parity suites, smoke, bench line, then (same command already exited 0) the ncu launch list and
# one full capture of the dominant kernel.  Usage: bash scripts/gpu_round.sh <tag> [skip-ncu]
tag=${1:-r1}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh encode network search dropin; trc=$?
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; grep -v "^\[W" gpurun_out/bench_$tag.err | tail -n 5
if [ "$2" != "skip-ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-sims 2"
  $CMD > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
  echo "ncu launches rc=$?"
  $CMD > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:tower_tcgen05|mcts_step|encode_planes" -s 9 -c 3 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
  echo "ncu full rc=$?"; tail -n 3 gpurun_out/ncu_full_$tag.log
fi
exit $trc
