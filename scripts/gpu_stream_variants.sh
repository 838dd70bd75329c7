#!/bin/bash
mkdir -p gpurun_out
for mode in hilo hilo@5 hilo@10 hilo@15 bf16; do
  timeout 600 python bench.py --stream $mode --steps 20 --warmup 5 --no-cpu-baseline --e2e-sims 20 > gpurun_out/bench_variant_$mode.json 2> gpurun_out/bench_variant.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_variant_$mode.json"))
print("$mode", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "tower us", round(d["kernels"]["tower"]["avg_us"],1), "TF", round(d["roofline"]["achieved"],1), "parity dlogit", round(d["parity"]["max_dlogit"],4), "dvalue", round(d["parity"]["max_dvalue"],4), d["parity"]["holds"])
PY
done
