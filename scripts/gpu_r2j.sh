#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_network.jsonl
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_dropin.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_network.log 2>&1; echo "== network+dropin rc=$?"; grep -v PARITY gpurun_out/test_network.log | tail -n 12
grep "PARITY.*oracle" gpurun_out/test_network.log | cut -c1-330
for net in 6x128 20x256 4x48; do
timeout 600 python bench.py --workload leaf_eval --net $net --steps 200 > gpurun_out/bench_leaf_$net.json 2> gpurun_out/bench_leaf_$net.err; echo "leaf $net rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_leaf_$net.json"))
print("$net value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["fraction_of_device_rate"],3), {k: round(v["avg_us"],1) for k,v in d["kernels"].items()})
PY
done
