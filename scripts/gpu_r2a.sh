#!/bin/bash
# round 2, first GPU session: the changed tower kernel first (short timeout), then the suites, smoke, bench in both stream modes
mkdir -p gpurun_out
rm -f gpurun_out/parity_network.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu -s -p no:cacheprovider -x > gpurun_out/test_network.log 2>&1; echo "== network rc=$?"; tail -n 30 gpurun_out/test_network.log
grep PARITY gpurun_out/test_network.log
for f in encode search dropin train; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f rc=$?"; tail -n 12 gpurun_out/test_$f.log
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; cat gpurun_out/bench_r2a.json; grep -v "^\[W" gpurun_out/bench_r2a.err | tail -n 8
timeout 600 python bench.py --stream bf16 --no-cpu-baseline > gpurun_out/bench_r2a_bf16stream.json 2> gpurun_out/bench_r2a_bf16stream.err; echo "bench bf16 rc=$?"; cat gpurun_out/bench_r2a_bf16stream.json
