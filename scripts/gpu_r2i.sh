#!/bin/bash
# full -m gpu suite + smoke + the new bench workloads at reduced size
mkdir -p gpurun_out
rm -f gpurun_out/parity_network.jsonl
for f in network encode search dropin train; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f rc=$?"; tail -n 4 gpurun_out/test_$f.log
done
timeout 600 python -m pytest tests/test_reference_suite.py tests/test_keras_io.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py --workload puzzles --positions 2048 --sims 100 --agreement-positions 16 > gpurun_out/bench_puzzles_small.json 2> gpurun_out/bench_puzzles_small.err; echo "puzzles rc=$?"; cat gpurun_out/bench_puzzles_small.json | cut -c1-1500; tail -n 3 gpurun_out/bench_puzzles_small.err
timeout 600 python bench.py --workload selfplay --games 512 --sims 100 --steps 2 > gpurun_out/bench_selfplay_small.json 2> gpurun_out/bench_selfplay_small.err; echo "selfplay rc=$?"; cat gpurun_out/bench_selfplay_small.json | cut -c1-1500; tail -n 3 gpurun_out/bench_selfplay_small.err
timeout 600 python bench.py --workload selfplay --games 64 --sims 100 --steps 2 > gpurun_out/bench_selfplay64_small.json 2> gpurun_out/bench_selfplay64_small.err; echo "selfplay64 rc=$?"; cat gpurun_out/bench_selfplay64_small.json | cut -c1-1500; tail -n 3 gpurun_out/bench_selfplay64_small.err
