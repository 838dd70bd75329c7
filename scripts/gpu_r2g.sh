#!/bin/bash
# regime study: long timed regions (power-capped steady state) in both stream modes, and a batch that stays in L2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,power.max_limit,power.default_limit,enforced.power.limit,clocks.max.sm --format=csv
for cfg in "hilo 1024 1500" "bf16 1024 1500" "hilo 256 1500" "bf16 256 1500" "hilo 512 1500"; do
  set -- $cfg
  timeout 600 python bench.py --stream $1 --trees $2 --steps $3 --no-cpu-baseline --no-parity --e2e-sims 4 > gpurun_out/bench_r2g_$1_$2.json 2> gpurun_out/bench_r2g_$1_$2.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2g_$1_$2.json"))
print("$cfg", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "tower us", round(d["kernels"]["tower"]["avg_us"],1), "achieved TF", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"], d["clocks"]["reasons"], "samples", d["clocks"]["samples"])
PY
done
