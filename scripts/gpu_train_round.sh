#!/bin/bash
# Training-step GPU session: parity suite, bench line, then (same command already exited 0) the ncu launch list and one
# full capture of the wgrad / BN kernels.  Usage: bash scripts/gpu_train_round.sh <tag> [skip-ncu]
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_train.log 2>&1; trc=$?
echo "== test_gpu_train rc=$trc"; grep -E "^train step|passed|failed" gpurun_out/test_train.log
python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train_$tag.json 2> gpurun_out/bench_train_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_train_$tag.json; grep -v "^\[W" gpurun_out/bench_train_$tag.err | tail -n 5
if [ "$2" != "skip-ncu" ]; then
  CMD="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_train_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 900 --csv --log-file gpurun_out/launches_train_$tag.csv $CMD > gpurun_out/ncu_launches_train_$tag.log 2>&1
  echo "ncu launches rc=$?"
  ncu --set full --clock-control none --import-source on -k "regex:wgrad_tcgen05|bn_bwd_apply|bn_stats|bn_apply_kernel|head_conv_bwd|sgemm" -s 60 -c 6 -f -o gpurun_out/prof_train_$tag $CMD > gpurun_out/ncu_full_train_$tag.log 2>&1
  echo "ncu full rc=$?"; tail -n 3 gpurun_out/ncu_full_train_$tag.log
fi
exit $trc
