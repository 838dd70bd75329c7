#!/bin/bash
# BASELINE configs 4 and 3 at their stated sizes on the GPUs of this box (N = number of GPUs visible)
N=${1:-1}
mkdir -p gpurun_out/final
O=gpurun_out/final
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; fi
$RUN bench.py --gpus $N --workload puzzles --positions 10000 --sims 400 --agreement-positions 32 > $O/bench_puzzles_10000x400_n$N.json 2> $O/bench_puzzles_n$N.err; echo "puzzles rc=$?"; cut -c1-1400 $O/bench_puzzles_10000x400_n$N.json; tail -n 2 $O/bench_puzzles_n$N.err
$RUN bench.py --gpus $N --workload selfplay --games 512 --sims 800 --steps 3 > $O/bench_selfplay_512x800_n$N.json 2> $O/bench_selfplay_n$N.err; echo "selfplay rc=$?"; cut -c1-1400 $O/bench_selfplay_512x800_n$N.json; tail -n 2 $O/bench_selfplay_n$N.err
if [ "$N" != "1" ]; then
  $RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_search_20x256_n$N.json 2> $O/bench_search_n$N.err; echo "search rc=$?"; cut -c1-600 $O/bench_search_20x256_n$N.json
  $RUN bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/bench_reference_n$N.json 2>/dev/null; echo "reference arm rc=$?"; cut -c1-200 $O/bench_reference_n$N.json
fi
