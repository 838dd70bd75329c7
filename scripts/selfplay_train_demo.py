"""One process, one GPU: the reference's actor / trainer cycle (train.py:21-52 play_game -> buffer, train.py:150-164 model.fit,
train.py:88-101 network hand-over) on the B200 path, at toy scale by default:

    python scripts/selfplay_train_demo.py [--net 4x48] [--games 8] [--iters 2] [--batch 32]

Self-play games run as one batch of concurrent trees (selfplay.play_games), their samples go to the replay buffer in the
reference's format, the Trainer takes SGD steps on sampled batches, and the updated weights are loaded back into the
inference network (no TF-TRT conversion step).  Prints one line per iteration."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(net="4x48", games=8, iters=2, batch=32, steps_per_iter=4, sims=(6, 12), seed=0, max_game_length=24, log=print):
    from chessbot_b200 import selfplay
    from chessbot_b200.config import Config
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    from chessbot_b200.train import Trainer, sample_buffer, scheduler
    blocks, filters = (int(x) for x in net.split("x"))
    config = Config()
    config.conv_filters, config.num_residual_blocks = filters, blocks
    config.num_mcts_sims, config.max_game_length, config.batch_size = tuple(sims), max_game_length, batch
    weights = keras_init_weights(filters, blocks, seed=seed)
    trainer = Trainer(weights, filters, blocks, batch, config=config)
    network = DeviceNetwork(weights, filters, blocks, max_batch=max(games, 8))
    buffer, history, lr = [], [], config.learning_rate[0]
    for it in range(iters):
        t0 = time.perf_counter()
        rngs = [np.random.default_rng(seed * 1000 + it * games + g) for g in range(games)]
        for result in selfplay.play_games(network, games, config, rngs):
            buffer.extend(selfplay.buffer_entries(result))
        buffer = buffer[-config.buffer_size:]                                  # FIFO trim (train.py:165-166)
        t1 = time.perf_counter()
        logs = None
        if len(buffer) >= batch:
            for s in range(steps_per_iter):
                x, (yv, yp) = sample_buffer(buffer, batch, np.random.default_rng(seed + 17 * it + s))
                lr = scheduler(trainer.steps, lr, config)
                logs = trainer.fit(x, {"value_head": yv, "policy_head": yp}, lr, shuffle=False)
            network = DeviceNetwork(trainer.get_weights(), filters, blocks, max_batch=max(games, 8))   # hand-over to the actors
        history.append((len(buffer), logs))
        log(f"iter {it}: buffer {len(buffer)} samples, self-play {t1 - t0:.2f} s, train {time.perf_counter() - t1:.2f} s, losses {logs}")
    return history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="4x48")
    ap.add_argument("--games", type=int, default=8)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    run(a.net, a.games, a.iters, a.batch)
