#!/bin/bash
# Runs the -m gpu parity suites one file at a time (a hang in one kernel must not hide the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in ${@:-encode network search dropin train}; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -s -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  r=$?
  echo "== test_gpu_$f rc=$r"; grep -v "^PARITY" gpurun_out/test_$f.log | tail -n 6
  [ $r -ne 0 ] && rc=$r
done
timeout 600 python -m pytest tests/test_reference_suite.py tests/test_keras_io.py -q -m gpu -p no:cacheprovider > gpurun_out/test_misc.log 2>&1; r=$?
echo "== reference suite + keras_io rc=$r"; tail -n 3 gpurun_out/test_misc.log
[ $r -ne 0 ] && rc=$r
exit $rc
