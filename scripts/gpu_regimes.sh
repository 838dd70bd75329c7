#!/bin/bash
# conv parity, then long timed regions (power-capped steady state) in both stream modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu -p no:cacheprovider -k "conv_layers or benched_shape or batch_invariant" > gpurun_out/test_network_b.log 2>&1; echo "== network rc=$?"; tail -n 3 gpurun_out/test_network_b.log
for cfg in "hilo 1024 1500" "bf16 1024 1500" "hilo 1024 100" "bf16 1024 100"; do
  set -- $cfg
  timeout 600 python bench.py --stream $1 --trees $2 --steps $3 --no-cpu-baseline --no-parity --e2e-sims 4 > gpurun_out/bench_$TAG_$1_$2_$3.json 2> gpurun_out/bench_$TAG_$1_$2_$3.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG_$1_$2_$3.json"))
print("$cfg", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "tower us", round(d["kernels"]["tower"]["avg_us"],1), "achieved TF", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"], d["clocks"]["reasons"], "samples", d["clocks"]["samples"])
PY
done
