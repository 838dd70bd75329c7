#!/bin/bash
# the driver's own invocation (20 timed steps, 5 warm-up: burst regime) in both stream modes, three times each
mkdir -p gpurun_out
for rep in 1 2 3; do
for mode in hilo bf16; do
  timeout 600 python bench.py --stream $mode --steps 20 --warmup 5 --no-cpu-baseline --no-parity --e2e-sims 20 > gpurun_out/bench_burst_$mode.json 2> gpurun_out/bench_burst_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_burst_$mode.json"))
print("$mode rep $rep", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "tower us", round(d["kernels"]["tower"]["avg_us"],1), "TF", round(d["roofline"]["achieved"],1), "mcts us", round(d["kernels"]["mcts_step"]["avg_us"],1), "e2e", round(d["e2e"]["value"]))
PY
done
done
