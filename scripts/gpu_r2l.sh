#!/bin/bash
# tower change check: conv / forward parity + training step (one-layer launches) + burst and sustained benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -p no:cacheprovider > gpurun_out/test_network.log 2>&1; echo "== network rc=$?"; tail -n 3 gpurun_out/test_network.log
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -p no:cacheprovider -k "autograd or wgrad or config5" > gpurun_out/test_train.log 2>&1; echo "== train rc=$?"; tail -n 3 gpurun_out/test_train.log
timeout 900 python -m pytest tests/test_gpu_search.py -q -m gpu -p no:cacheprovider > gpurun_out/test_search.log 2>&1; echo "== search rc=$?"; tail -n 3 gpurun_out/test_search.log
bash scripts/gpu_r2k.sh 2>&1 | tail -4
TAG=r2l bash scripts/gpu_r2h.sh 2>&1 | grep "1024 1500"
