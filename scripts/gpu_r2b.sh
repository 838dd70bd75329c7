#!/bin/bash
# quick A/B of the tower kernel: conv parity + short bench in both residual-stream modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu -s -p no:cacheprovider -k "conv_layers or benched_shape or batch_invariant" > gpurun_out/test_network_b.log 2>&1; echo "== network rc=$?"; tail -n 5 gpurun_out/test_network_b.log; grep PARITY gpurun_out/test_network_b.log
for mode in hilo bf16; do
  timeout 600 python bench.py --stream $mode --no-cpu-baseline --no-parity --steps 100 --e2e-sims 20 > gpurun_out/bench_$1_$mode.json 2> gpurun_out/bench_$1_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1_$mode.json"))
print("$mode", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "tower us", round(d["kernels"]["tower"]["avg_us"],1), "achieved", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]), d["e2e"]["fraction_of_device_rate"])
PY
done
