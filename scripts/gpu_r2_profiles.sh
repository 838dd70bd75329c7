#!/bin/bash
# final-build evidence for profiles/: bench lines (default = driver-style + long), ncu launch list and full capture of the same command
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,memory.total --format=csv > $O/smi.txt
python bench.py --steps 20 --warmup 5 > $O/bench_search_20x256_driver_style.json 2> $O/bench_driver_style.err; echo "driver-style rc=$?"
python bench.py > $O/bench_search_20x256_default.json 2> $O/bench_default.err; echo "default rc=$?"
python bench.py --stream bf16 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_search_20x256_bf16stream_driver_style.json 2>/dev/null; echo "bf16 stream rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
python bench.py --net 4x48 --no-cpu-baseline > $O/bench_search_4x48.json 2>/dev/null; echo "4x48 rc=$?"
python bench.py --net 6x128 --no-cpu-baseline > $O/bench_search_6x128.json 2>/dev/null; echo "6x128 rc=$?"
python bench.py --workload leaf_eval --net 6x128 > $O/bench_leaf_eval_6x128.json 2>/dev/null; echo "leaf rc=$?"
# config 1: one game, 100 sims, the reference's default 4x48 net
python - > $O/config1_single_game_latency.json 2>$O/config1.err <<'PY'
import json, time, sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from chessbot_b200 import mcts
from chessbot_b200.config import Config
from chessbot_b200.game import Game
from chessbot_b200.model import DeviceNetwork, keras_init_weights
out = {}
for net_s in ("4x48", "20x256"):
    f, b = bench.parse_net(net_s)
    net = DeviceNetwork(bench.trained_like_bn(keras_init_weights(f, b, seed=0), f, b), f, b, max_batch=64)
    for leaves in (1, 8):
        mcts.run_mcts(Game(Config()), net, 100, 30, True, None, leaves_per_tree=leaves)
        torch.cuda.synchronize(); t0 = time.perf_counter(); n = 10
        for _ in range(n):
            mcts.run_mcts(Game(Config()), net, 100, 30, True, None, leaves_per_tree=leaves)
        torch.cuda.synchronize()
        out[f"{net_s} leaves_per_tree={leaves}"] = {"ms_per_100_sim_search": 1e3 * (time.perf_counter() - t0) / n}
print(json.dumps({"what": "BASELINE config 1: mcts.run_mcts from the start position, 100 sims, batch 1 (wall clock per call, host objects in / move out)", "results": out}))
PY
echo "config1 rc=$?"; cat $O/config1_single_game_latency.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-sims 2"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_bench_20x256_b1024.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:tower_tcgen05|mcts_step|encode_planes|head_features" -s 12 -c 4 -f -o $O/prof_final $CMD > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_final.ncu-rep --page raw --csv > $O/ncu_full_raw_encode_tower_heads_mcts.csv 2>/dev/null
cuobjdump -sass -fun '_ZN2cb20tower_tcgen05_kernelILb0EEEvNS_11TowerParamsE' chessbot_b200/libchessbot_b200.so 2>/dev/null | grep -E "UTCHMMA|UTMALDG|UTCBAR|LDTM|USETMAXREG|SYNCS" | sort | uniq -c | sort -rn > $O/sass_tower_tcgen05_mnemonics.txt
cuobjdump -sass -fun 'wgrad_tcgen05_kernel' chessbot_b200/libchessbot_b200.so 2>/dev/null | grep -E "UTCHMMA|UTMALDG|UTCBAR|LDTM|SYNCS" | sort | uniq -c | sort -rn > $O/sass_wgrad_tcgen05_mnemonics.txt
ls -la $O
