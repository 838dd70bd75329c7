#!/usr/bin/env python
"""bench.py — the self-play hot path (batched MCTS leaf evaluation + tree update) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--net 20x256] [--trees 1024] [--impl reference]

A "step" is ONE pass of the hot path over one batch: for every one of `trees` concurrent games the
pending leaf is encoded (K1), evaluated by the residual tower + heads (K2-K4), expanded (K5: legal-move
gather, exp, normalise), its value backed up (K8), and the next leaf is selected with the device chess
rules (K6/K7).  value = NN position evaluations per second over all ranks (BASELINE.json metric: "MCTS
sims/sec + NN position evals/sec"); sims/s (>= evals/s: terminal leaves need no network) is reported
beside it.  Inputs: seeded random-playout positions, glorot random-init kernels with trained-like BN statistics (no data or weights ship
with the reference).  Multi-GPU: games are sharded by rank, no data-path collective.

`--impl reference` times the reference's CPU path (its mcts.py/gameimage.py/actionspace.py logic as
restated in oracle/, python-chess and TensorFlow replaced by the oracle shims; torch-fp32 CPU network)
on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MCTS sims/sec + NN position evals/sec (value = NN position evals/s of batched MCTS; sims/s beside it)"
UNIT = "evals/s"


def parse_net(s):
    blocks, filters = s.lower().split("x")
    return int(filters), int(blocks)


def flops_per_position(filters, blocks, in_channels=119):
    C = filters
    return 2 * 64 * 9 * C * (in_channels + 2 * blocks * C) + 2 * 64 * C * 3 + 2 * 64 * C + 2 * C + 2 * 128 * 4672


def conv_flops_per_position(filters, blocks, in_channels=119):
    return 2 * 64 * 9 * filters * (in_channels + 2 * blocks * filters)


def measured_peaks(timed_s=None):
    """(bf16 TFLOP/s peak to compare with, HBM GB/s, source, {"burst": .., "sustained": ..}).  MEASURED_PEAKS.json holds two bf16
    figures: the burst one (a kernel timed alone / a timed region under ~1 s, before the power cap pulls the clock down) and the
    sustained one (a kernel inside a long step).  The one used follows the DURATION of the timed region."""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        both = {"burst": float(p["bf16_tflops"]), "sustained": float(p["bf16_tflops_sustained"])}
        hbm, src = float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        both, hbm, src = {"burst": 1650.0, "sustained": 1400.0}, 6650.0, "fallback (B200_PROFILING.md)"
    kind = "sustained" if (timed_s is None or timed_s >= 1.0) else "burst"
    return both[kind], hbm, f"{src}, {kind} bf16 figure (timed region {timed_s if timed_s is None else round(timed_s, 3)} s)", both


def profiled_traffic(kernel_substr, pattern="r*_ncu_full_raw*tower*.csv"):
    """DRAM read + write bytes per launch of a kernel from the newest committed `ncu --set full` raw page under profiles/ (a number
    taken under the profiler, on an earlier run of the same kernel: labelled as such, never measured inside this run)."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if kernel_substr in r[ik]:
                    return float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]], os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


# ------------------------------------------------------------------------------------------ inputs
def trained_like_bn(w, filters, blocks, in_channels=119):
    """Overwrites the BatchNorm moving statistics of a Keras-layout weights dict (glorot kernels stay) with
    the values a trained network would hold for them, from a closed-form variance recurrence: the stem
    normalises the raw integer planes (move count ~80, halfmove clock ~50), every BN sees unit-variance
    pre-activations.  Logits come out O(1) (|logit| max 1..2, checked against the torch-fp32 oracle).
    Fresh Keras BN (mean 0 / var 1) lets a 20x256 tower reach |logit| ~ 95: np.exp overflows, priors
    become NaN and every search degenerates — not a workload."""
    C = filters

    def bn(c, var):
        return np.stack([np.ones(c), np.zeros(c), np.zeros(c), np.full(c, var)]).astype(np.float32)

    w = dict(w)
    w["stem.bn"] = bn(C, 0.9 * 2.0 * in_channels / (in_channels + C) * (80.0 ** 2 + 50.0 ** 2 + 30.0) / in_channels)
    m = 0.6                                     # second moment of LeakyReLU(0.3) of a unit normal ~ 0.545 (+ mean)
    for i in range(blocks):
        w[f"res{i}.bn1"] = bn(C, 0.85 * m)      # 0.85: border taps of the 3x3 "same" convolution
        w[f"res{i}.bn2"] = bn(C, 0.85 * 0.545)
        m = 0.6 * (m / 0.6 + 1.0)               # x + BN(y): variances add
    w["value.bn"] = bn(1, m * 2 * C / (C + 1))
    w["policy.bn"] = bn(2, m * 2 * C / (C + 2))
    return w


def random_positions_device(n, seed, min_plies=8, max_plies=160):
    """seeded random playouts with the product's own host rules; full move stacks."""
    import random
    from chessbot_b200.board import Board
    rnd = random.Random(seed)
    boards = []
    while len(boards) < n:
        b = Board()
        ok = True
        for _ in range(rnd.randint(min_plies, max_plies)):
            if b.is_game_over(claim_draw=True):
                ok = False
                break
            ms = list(b.legal_moves)
            b.push(rnd.choice(ms))
        if ok and not b.is_game_over(claim_draw=True):
            boards.append(b)
    return boards


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------ CPU legs
def _oracle_setup(filters, blocks, seed):
    import torch
    from oracle import ref_model
    w = trained_like_bn(ref_model.init_weights(filters, blocks, seed=seed), filters, blocks)
    return ref_model.as_network(w, "cpu")


def _oracle_search(args):
    """One bounded search with the reference-path oracle: returns (evals, sims)."""
    moves, sims, filters, blocks, threads = args
    import torch
    torch.set_num_threads(threads)
    from oracle import pychess_shim as chess
    from oracle import ref_mcts
    from oracle.ref_game import Game, OracleConfig
    global _ORACLE_NET
    try:
        net = _ORACLE_NET
    except NameError:
        net = _ORACLE_NET = _oracle_setup(filters, blocks, 0)
    b = chess.Board()
    for m in moves:
        b.push_uci(m)
    stats = {}
    ref_mcts.run_mcts(Game(OracleConfig(), board=b), net, sims, 30, True, None, rng=np.random.default_rng(1), stats=stats)
    return stats["evals"], sims


def _oracle_positions(n, seed):
    import random
    from oracle import pychess_shim as chess
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        b = chess.Board()
        ok = True
        for _ in range(rnd.randint(8, 160)):
            if b.is_game_over(claim_draw=True):
                ok = False
                break
            b.push(rnd.choice(list(b.legal_moves)))
        if ok and not b.is_game_over(claim_draw=True):
            out.append([m.uci() for m in b.move_stack])
    return out


def cpu_baseline_leg(filters, blocks, budget_s=12.0, sims=16):
    """Reference-path oracle, single process, torch CPU threads as torch picks them; ~budget_s of work."""
    import torch
    threads = torch.get_num_threads()
    pos = _oracle_positions(8, 123)
    t0 = time.perf_counter()
    evals = tot_sims = searches = 0
    while time.perf_counter() - t0 < budget_s:
        e, s = _oracle_search((pos[searches % len(pos)], sims, filters, blocks, threads))
        evals += e; tot_sims += s; searches += 1
    dt = time.perf_counter() - t0
    return {"value": evals / dt, "unit": UNIT, "sims_per_s": tot_sims / dt, "cores": threads, "kind": "port",
            "sample": f"{searches} searches x {sims} sims from seeded random-playout positions, {blocks}x{filters} torch-fp32 CPU network, "
                      f"1 process, {dt:.1f} s (python-chess/TensorFlow replaced by the oracle shims)"}


def parity_leg(eng, net, records, n):
    """Checker leg, outside every timed region (the oracle is test infrastructure: it checks, it is never what is measured): the
    network and the positions this run timed, device logits / values against the torch-fp32 oracle (oracle/ref_model.py, TF32 off,
    on the GPU through torch).  north_star tolerance: 2e-2 abs."""
    import torch
    from oracle import ref_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net.ensure_loaded()
    pos, fi = eng.upload_records(records)
    act, i16 = eng.encode(pos, fi, want_act=True, want_i16=True)
    value, _, lf = eng.forward(act, n, want_bf16=False, want_f32=True)
    torch.cuda.synchronize()
    v_ref, p_ref = ref_model.forward(net.weights, i16.cpu().numpy(), device="cuda")
    dp, dv = np.abs(lf.cpu().numpy() - p_ref), np.abs(value.cpu().numpy() - v_ref)
    return {"n": int(n), "max_dlogit": float(dp.max()), "p999_dlogit": float(np.quantile(dp, 0.999)), "rms_dlogit": float(np.sqrt(np.mean(dp ** 2))),
            "max_dvalue": float(dv.max()), "top1_agree": float((lf.cpu().numpy().argmax(1) == p_ref.argmax(1)).mean()), "tolerance_abs": 2e-2,
            "holds": bool(dp.max() <= 2e-2 and dv.max() <= 2e-2),
            "against": "oracle/ref_model.py (torch fp32, TF32 off) on the positions and weights of the timed run; residual stream "
                       + ("bf16" if not net.stream_hi_lo else "bf16 hi + lo" if net.stream_hi_lo is True else f"hi + lo from block {net.stream_hi_lo - 1}")}


def run_reference_arm(args):
    """`--impl reference`: the reference CPU path on all host cores: P independent processes (the
    reference's own scaling axis is num_actors processes, train.py:173,198-201), each step = P bounded
    searches."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    filters, blocks = parse_net(args.net)
    if args.workload == "train":
        return run_reference_train_arm(args, filters, blocks)
    P = min(os.cpu_count() or 1, 64)
    sims = 8
    pos = _oracle_positions(P, 321)
    ctx = mp.get_context("spawn")
    with ctx.Pool(P) as pool:
        job = [(pos[i], sims, filters, blocks, 1) for i in range(P)]
        for _ in range(max(args.warmup, 1)):
            pool.map(_oracle_search, job)
        t0 = time.perf_counter()
        evals = tot = 0
        for _ in range(args.steps):
            for e, s in pool.map(_oracle_search, job):
                evals += e; tot += s
        dt = time.perf_counter() - t0
    value = evals / dt
    sample = (f"{args.steps} steps x {P} processes x one {sims}-sim search each ({blocks}x{filters} torch-fp32 CPU network, 1 thread per "
              f"process), seeded random-playout positions; python-chess/TensorFlow replaced by the oracle shims")
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "sims_per_s": tot / dt, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batched MCTS leaf evaluation + tree update, {blocks}x{filters} net (reference CPU path, batch 1 per process)",
                   "net": args.net, "processes": P,
                   "comparison": "a CPU PORT of the reference path (python-chess / TensorFlow replaced by the oracle shims, torch-fp32 network), bounded sample: "
                                 f"{sims}-sim searches in fp32 against the GPU arm's 100-sim searches in bf16 - evaluations per second are what is compared, "
                                 "not identical searches; the reference itself evaluates with TF-TRT on a GPU (108 sims/s on an RTX 2070 SUPER, main.ipynb)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from chessbot_b200.engine import EVAL_NET, Engine
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    from chessbot_b200 import mcts as cb_mcts
    from chessbot_b200.game import Game
    from chessbot_b200.config import Config

    filters, blocks = parse_net(args.net)
    trees = args.trees
    W, K = args.warmup, args.steps
    eng = Engine.get(local_rank)
    net = DeviceNetwork(trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks), filters, blocks,
                        max_batch=trees, device=local_rank,
                        stream_hi_lo=(args.stream == "hilo") if "@" not in args.stream else int(args.stream.split("@")[1]) + 1)
    boards = random_positions_device(trees, seed=1000 + rank)   # every rank searches its own shard of games
    records = [b.export_record() for b in boards]
    total_steps = W + 2 * K + 4
    pool_steps = max(total_steps, 720)              # also holds the same-regime device leg behind the e2e leg (<= 700 steps)
    eng.search_create(trees, max_nodes=trees * (pool_steps + 2) * 48, max_positions=trees * (pool_steps + 200),
                      max_sims=max(800, total_steps))
    rng = np.random.default_rng(7 + rank)
    noise = [rng.gamma(0.3, 1, len(b.legal_moves)) for b in boards]
    eng.search_begin(records, max(800, total_steps), noise=noise, eval_mode=EVAL_NET)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)               # started before the warm-up: nvidia-smi needs ~0.2 s to deliver its first sample
    sampler.start()
    eng.search_run(W)                                # warm-up steps (incl. the root expansion)
    barrier()
    c0 = eng.search_status()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    eng.search_run(K)                                # EXACTLY K timed steps
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    c1 = eng.search_status()
    evals, sims = c1["evals"] - c0["evals"], c1["sims"] - c0["sims"]
    if c1["errors"]:
        raise RuntimeError("search pool exhausted during the bench")

    # roofline leg: the same K steps again with CUDA events (on the launching stream) around every kernel launch; the pass is also
    # timed as a whole, so the kernels' shares refer to THIS pass (per-launch events cost it a few per cent against the timed one)
    nodes0 = int(eng.search_read_roots()["nodes"].sum())
    eng.profile(True)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pv0.record(stream)
    eng.search_run(K)
    pv1.record(stream)
    prof = eng.profile_read()
    prof_pass_ms = pv0.elapsed_time(pv1)
    eng.profile(False)
    c2 = eng.search_status()
    nodes1 = int(eng.search_read_roots()["nodes"].sum())
    conv_ms, conv_launches = prof["conv"]
    evals_prof = c2["evals"] - c1["evals"]                                           # boards the tower really evaluated in this pass (terminal
    conv_flops_launch_total = conv_flops_per_position(filters, blocks) * evals_prof   # leaves need none): algorithmic direct-conv FLOPs
    peak_tf, peak_gbs, peak_src, peak_both = measured_peaks(ms * 1e-3)
    achieved_tf = conv_flops_launch_total / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # HBM-bound kernels, algorithmic bytes as stated in DESIGN.md §4
    enc_bytes = 17152.0 * evals_prof                                            # 8 x 96 B frames read + 128 ch bf16 written
    # prior gather + node writes; select / backup; + the fused encode of the selected leaf and the head features of the expanded one
    mcts_bytes = (nodes1 - nodes0) * (512 + 32 + 2) + (c2["sims"] - c1["sims"]) * 2700.0 + evals_prof * (17152.0 + 768.0)
    kernels = {}
    for name, nbytes in (("encode", enc_bytes), ("mcts_step", mcts_bytes), ("heads", None)):
        ms_k, n_k = prof["heads" if name == "heads" else name]
        kernels[name] = {"launches": n_k, "avg_us": 1e3 * ms_k / max(n_k, 1), "share_of_step": ms_k / prof_pass_ms}
        if n_k == 0:
            kernels[name]["note"] = "fused into mcts_step_kernel in the search loop (stand-alone kernel only for the roots / cb_net_forward)"
            continue
        if nbytes:
            gbs = nbytes / (ms_k * 1e-3) / 1e9
            kernels[name].update({"bound": "hbm", "achieved_gbs": gbs, "peak_gbs": peak_gbs, "frac": gbs / peak_gbs})
    kernels["tower"] = {"launches": conv_launches, "avg_us": 1e3 * conv_ms / max(conv_launches, 1), "share_of_step": conv_ms / prof_pass_ms}
    traffic, traffic_src = profiled_traffic("tower_tcgen05_kernel") if (filters, blocks, trees) == (256, 20, 1024) else (None, None)

    # end-to-end leg: complete searches through the public API from host game objects: host rules export the game records, H2D,
    # e2e_sims+1 device steps, D2H of the root statistics, move choice.  (a) mcts.SearchStream, the call selfplay.play_games makes:
    # two groups of `trees` games alternate on the device, the host work of one group overlaps the search of the other;
    # (b) one mcts.run_mcts_batch call at a time (nothing overlaps).
    cfg = Config()
    e2e_sims = args.e2e_sims
    groups = [[Game(cfg, board=b) for b in boards], [Game(cfg, board=b) for b in random_positions_device(trees, seed=5000 + rank)]]
    rng_e = np.random.default_rng(3)

    def caller_share(res):
        # play_game reads the root's children (visit distribution, train.py:13-19,43-47) on its full searches only, a quarter of
        # the moves (config.num_mcts_sims_p); the other moves just take the move
        return sum(sum(c.N for c in root.children.values()) for _, root in res[::4])

    ss = cb_mcts.SearchStream(net, 2)
    for g in groups:                                     # warm both lanes (allocations, first-touch)
        ss.collect(ss.submit(g, e2e_sims, True, None, rng=rng_e))
    barrier()
    rounds, touched, e2e_evals_exact = 3, 0, 0
    t0 = time.perf_counter()
    tickets = [ss.submit(g, e2e_sims, True, None, rng=rng_e) for g in groups]
    for r in range(rounds):
        for gi, g in enumerate(groups):
            res = ss.collect_moves(tickets[gi], g, 30, rng_e)
            e2e_evals_exact += ss.last_counters["evals"]
            touched += caller_share(res)
            if r + 1 < rounds:
                tickets[gi] = ss.submit(g, e2e_sims, True, None, rng=rng_e)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    # the device-only rate in the SAME clock regime as the e2e leg (it runs ~1 s and sits at the power cap; the K timed steps of `value`
    # may be a burst): as many device steps as the e2e leg ran, inputs resident, CUDA events - right behind it, same thermal state
    eng.search_begin(records, max(800, total_steps), noise=noise, eval_mode=EVAL_NET)
    eng.search_run(W)
    c_a = eng.search_status()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    same_steps = min(2 * rounds * (e2e_sims + 1), 700)
    s0.record(stream)
    eng.search_run(same_steps)
    s1.record(stream)
    torch.cuda.synchronize()
    c_b = eng.search_status()
    if c_b["errors"]:
        raise RuntimeError("search pool exhausted during the same-regime device leg")
    device_same_regime = (c_b["evals"] - c_a["evals"]) / (s0.elapsed_time(s1) * 1e-3)
    # (b) single calls
    cb_mcts.run_mcts_batch(groups[0], net, e2e_sims, 30, True, None, rng=rng_e)
    barrier()
    t1 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        touched += caller_share(cb_mcts.run_mcts_batch(groups[0], net, e2e_sims, 30, True, None, rng=rng_e))
    torch.cuda.synchronize()
    single_dt = time.perf_counter() - t1
    single_evals = net.engine.search_status()["evals"] * reps
    n_children = sum(len(b.legal_moves) for b in boards)
    h2d = sum(r.nbytes for r in records) + trees * (80 + 4 + 4) + n_children * 8 + (trees + 1) * 4
    d2h = n_children * 22 + trees * 16 + 16

    if world > 1:
        t = torch.tensor([ms, e2e_dt, conv_ms, single_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([evals, sims, e2e_evals_exact, single_evals, device_same_regime], device="cuda", dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        ms, e2e_dt, conv_ms_max, single_dt = t.tolist()
        evals, sims, e2e_evals_exact, single_evals, device_same_regime = s.tolist()
    value = evals / (ms * 1e-3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(filters, blocks)
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_leg(eng, net, records, trees)

    if rank == 0:
        steps_per_search = e2e_sims + 1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "sims_per_s": sims / (ms * 1e-3),
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"batched MCTS leaf evaluation + tree update: {trees} concurrent games per GPU, {blocks}x{filters} "
                            f"ResNet (glorot kernels, trained-like BN statistics), one leaf per game per step (encode + forward + legal-move softmax + backup + select)",
                "net": args.net, "trees_per_gpu": trees, "parallelism": f"games sharded over {world} GPU(s), no collective",
                "residual_stream": {"hilo": "bf16 hi + lo", "bf16": "bf16"}.get(args.stream, args.stream),
                "l2": "working set per step (49 MB weights + 3-4 x 52 MB activations + node pool) exceeds the 126 MB L2; no flush",
                "fwd_gflop_per_position": flops_per_position(filters, blocks) / 1e9,
            },
            "roofline": {
                "bound": "tensor", "kernel": "tower_tcgen05_kernel (all 1+2R conv layers, one persistent launch)", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "frac_of_burst_peak": achieved_tf / peak_both["burst"], "frac_of_sustained_peak": achieved_tf / peak_both["sustained"],
                "peaks": peak_both, "peak_source": peak_src,
                "traffic": traffic, "traffic_source": (f"NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one tower launch of this workload in the committed "
                                                       f"ncu --set full capture {traffic_src}") if traffic_src else None,
                "how": f"algorithmic direct-conv FLOPs of {conv_launches} launches / their CUDA-event durations ({conv_ms:.2f} ms) in a second, per-launch-"
                       f"instrumented pass of the same {K} steps ({prof_pass_ms / K:.3f} ms per step in that pass, {ms / K:.3f} in the timed one)",
                "conv_share_of_step": conv_ms / prof_pass_ms if prof_pass_ms > 0 else None,
                "profiled_pass_ms_per_step": prof_pass_ms / K,
            },
            "e2e": {"value": e2e_evals_exact / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d / steps_per_search,
                    "d2h_bytes_per_step": d2h / steps_per_search,
                    "fraction_of_device_rate": (e2e_evals_exact / e2e_dt) / device_same_regime,
                    "device_rate_same_regime": device_same_regime,
                    "fraction_of_value": (e2e_evals_exact / e2e_dt) / value,
                    "regime_note": f"`value` is timed over {K} steps ({ms * 1e-3:.2f} s: a burst if short); the e2e leg runs {e2e_dt:.2f} s at the power cap, so "
                                   f"it is compared with the device-only rate over {same_steps} steps measured right behind it (device_rate_same_regime)",
                    "what": f"mcts.SearchStream (the call selfplay.play_games makes): 2 groups of {trees} games alternate, {rounds} searches of {e2e_sims} sims "
                            f"each; per search: host game records -> H2D -> {steps_per_search} steps -> packed D2H of the root statistics -> move choice for "
                            "every game, children read for every 4th (playout cap); wall clock from the first submit to the last move, incl. the "
                            f"un-overlapped first export and last move choice; bytes are per search / {steps_per_search} steps",
                    "single_call": {"value": single_evals / single_dt, "fraction_of_device_rate": (single_evals / single_dt) / device_same_regime,
                                    "what": f"{reps} x mcts.run_mcts_batch({trees} games, {e2e_sims} sims), one call at a time: nothing overlaps the host work"}},
            "kernels": kernels,
            "gpu_launches": K * 2,   # tower, mcts_step (heads and the next leaf's encode are fused into mcts_step)
            "clocks": clocks,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_leaf_eval_arm(args):
    """BASELINE config 2: batched leaf evaluation of `trees` synthetic positions outside a tree — encode + forward
    (dense 4672 logits) + legal-move mask / softmax — on one GPU per rank.  A step = one pass over the batch."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from chessbot_b200.engine import Engine
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    filters, blocks = parse_net(args.net)
    n, W, K = args.trees, args.warmup, args.steps
    eng = Engine.get(local_rank)
    DeviceNetwork(trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks), filters, blocks, max_batch=n, device=local_rank)
    records = [b.export_record() for b in random_positions_device(n, seed=1000 + rank)]
    pos, fi = eng.upload_records(records)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(W, 50)):
        eng.evaluate_positions(None, pos, fi)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        eng.evaluate_positions(None, pos, fi)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    eng.profile(True)
    for _ in range(K):
        eng.evaluate_positions(None, pos, fi)
    prof = eng.profile_read()
    eng.profile(False)
    from chessbot_b200.board import export_records
    boards_h = random_positions_device(n, seed=1000 + rank)
    flat_h, lens_h, _ = export_records(boards_h, flat=True)
    eng.evaluate_records((flat_h, lens_h))                    # warm (scratch allocation)
    barrier()
    t0 = time.perf_counter()
    reps = max(5, K // 5)
    for _ in range(reps):                                    # end to end: host game records -> pinned staging -> H2D -> kernels -> packed priors D2H
        host = eng.evaluate_records((flat_h, lens_h))
    e2e_dt = time.perf_counter() - t0
    h2d = flat_h.nbytes + n * 36
    d2h = int(n * 12 + 4 + host["priors"].size * 8)
    if world > 1:
        t = torch.tensor([ms, e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_dt = t.tolist()
    if rank == 0:
        peak_tf, peak_gbs, peak_src, peak_both = measured_peaks(ms * 1e-3)
        conv_ms = prof["conv"][0]
        tf = conv_flops_per_position(filters, blocks) * n * K / (conv_ms * 1e-3) / 1e12
        tot = sum(v[0] for v in prof.values())
        emit(json.dumps({
            "metric": METRIC, "value": world * n * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"batched leaf eval (BASELINE config 2): {blocks}x{filters} ResNet, {n} synthetic positions per GPU, encode + forward + "
                                   "legal-move mask/softmax (policy Dense evaluated for the legal moves only)",
                       "net": args.net, "positions_per_gpu": n, "l2": "inputs resident; activations 3 x %.0f MB per pass" % (n * 32 * filters * 1.6 / 1e6)},
            "roofline": {"bound": "tensor", "kernel": "tower_tcgen05_kernel", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "frac_of_burst_peak": tf / peak_both["burst"], "frac_of_sustained_peak": tf / peak_both["sustained"], "traffic": None, "peak_source": peak_src},
            "kernels": {k: {"launches": v[1], "avg_us": 1e3 * v[0] / max(v[1], 1), "share_of_step": v[0] / tot} for k, v in prof.items() if v[1]},
            "e2e": {"value": world * n * reps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "fraction_of_device_rate": (world * n * reps / e2e_dt) / (world * n * K / (ms * 1e-3)),
                    "what": "Engine.evaluate_records (cb_leaf_eval_host_sync): host game records -> pinned staging -> H2D -> encode, tower, heads, "
                            "legal-move priors, pack -> D2H of value + CSR priors / moves / actions"},
            "gpu_launches": K * 4, "clocks": clocks, "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return world, rank, local_rank, barrier


def _reduce(world, maxes, sums):
    import torch
    import torch.distributed as dist
    if world > 1:
        t = torch.tensor(maxes, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor(sums, device="cuda", dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        return t.tolist(), u.tolist()
    return list(maxes), list(sums)


def _oracle_top_moves(boards, weights, sims):
    """Checker leg (untimed): the reference path's move for these positions — oracle/ref_mcts.py over the torch-fp32 network on the
    GPU (TF32 off), no root noise, most-visited move with ties to the first (like the device side here)."""
    import torch
    from oracle import pychess_shim as ochess, ref_mcts, ref_model
    from oracle.ref_game import Game as OGame, OracleConfig
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    class FirstTie:
        def choice(self, x):
            return x[0]
    net = ref_model.as_network(weights, "cuda")
    out = []
    for b in boards:
        ob = ochess.Board()
        for m in b.move_stack:
            ob.push_uci(m.uci())
        out.append(ref_mcts.run_mcts(OGame(OracleConfig(), board=ob), net, sims, 0, False, None, rng=FirstTie())[0])
    return out


def run_puzzles_arm(args):
    """BASELINE config 4 (puzzles.py:83-86,194-215; eval_trt.py:31-39): a sweep of `--positions` synthetic puzzle positions,
    `--sims` simulations each, no root noise, the most visited move is the answer.  Positions are sharded over the ranks; every
    rank feeds them through selfplay.Bot-style batches of `--trees` positions, two batches alternating on the device
    (mcts.SearchStream).  A step = one batch.  Top-move agreement against the reference path on a seeded subset (untimed)."""
    import torch
    from chessbot_b200 import mcts as cb_mcts
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    world, rank, local_rank, barrier = _dist_setup()
    filters, blocks = parse_net(args.net)
    trees, sims = args.trees, args.sims
    n_total = args.positions
    n_mine = len(range(rank, n_total, world))
    weights = trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks)
    weights["policy.fc"] = weights["policy.fc"] * 4.0        # a sharper policy than random init: searches get a favourite (a trained net's regime)
    net = DeviceNetwork(weights, filters, blocks, max_batch=trees, device=local_rank)
    boards = random_positions_device(n_mine, seed=7000 + rank, min_plies=8, max_plies=120)
    cfg = Config()
    batches = [[Game(cfg, board=b) for b in boards[i:i + trees]] for i in range(0, n_mine, trees)]
    class FirstTie:
        def choice(self, x):
            return x[0]
    rng = FirstTie()
    ss = cb_mcts.SearchStream(net, 2)
    for g in batches[:2]:
        ss.collect(ss.submit(g[:64], 8, False, None))       # warm both lanes
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    moves, evals, nsims, tickets = [], 0, 0, []
    for bi, g in enumerate(batches):
        tickets.append((ss.submit(g, sims, False, None), g))
        if len(tickets) == 2:
            tk, gg = tickets.pop(0)
            moves += [m for m, _ in ss.collect_moves(tk, gg, 0, rng)]
            evals += ss.last_counters["evals"]; nsims += ss.last_counters["sims"]
    for tk, gg in tickets:
        moves += [m for m, _ in ss.collect_moves(tk, gg, 0, rng)]
        evals += ss.last_counters["evals"]; nsims += ss.last_counters["sims"]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    agree = None
    if rank == 0 and not args.no_parity:
        k = min(args.agreement_positions, n_mine)
        ref = _oracle_top_moves(boards[:k], weights, sims)
        same = sum(a == b for a, b in zip(moves[:k], ref))
        agree = {"positions": k, "top_move_agreement": same / k, "sims": sims,
                 "against": "oracle/ref_mcts.py + torch-fp32 network (the reference path), same positions, no root noise"}
    (dt,), (evals, nsims, n_done) = _reduce(world, [dt], [evals, nsims, len(moves)])
    if rank == 0:
        emit(json.dumps({
            "metric": METRIC, "value": evals / dt, "unit": UNIT, "sims_per_s": nsims / dt, "positions_per_s": n_done / dt, "n_gpus": world,
            "steps": (n_total + trees * world - 1) // (trees * world), "warmup": 1, "ms_per_step": 1e3 * dt / max(1, len(batches)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"puzzles.py eval sweep (BASELINE config 4): {n_total} synthetic puzzle positions x {sims} sims, {blocks}x{filters} net, "
                                   f"batches of {trees} positions per GPU through mcts.SearchStream, most-visited move as the answer",
                       "net": args.net, "positions": n_total, "sims": sims, "trees_per_gpu": trees, "timing": "wall clock, host objects in / moves out (end to end)"},
            "agreement": agree, "e2e": {"value": evals / dt, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                        "what": "the whole sweep IS end to end: host boards -> records -> searches -> moves"},
            "gpu_launches": int(nsims / max(1, trees)) * 4, "clocks": clocks, "cpu_baseline": None}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_selfplay_arm(args):
    """BASELINE config 3 read literally (train.py:21-52): `--games` (512) concurrent self-play games in TOTAL, `--sims` (800) simulations per
    move, sharded by game over the ranks.  With fewer games per GPU the evaluation batch is kept at >= 512 positions by selecting
    several leaves per tree per step under virtual loss (leaves_per_tree = 512 / games per GPU).  A step = one move of every game (one
    complete search per game + move choice + the moves played).  Strong scaling: total work is fixed."""
    import torch
    from chessbot_b200 import mcts as cb_mcts
    from chessbot_b200.config import Config
    from chessbot_b200.game import Game
    from chessbot_b200.model import DeviceNetwork, keras_init_weights
    world, rank, local_rank, barrier = _dist_setup()
    filters, blocks = parse_net(args.net)
    n_total, sims = args.games, args.sims
    mine = list(range(rank, n_total, world))
    leaves = args.leaves if args.leaves > 0 else max(1, min(16, 512 // max(1, len(mine))))
    weights = trained_like_bn(keras_init_weights(filters, blocks, seed=0), filters, blocks)
    net = DeviceNetwork(weights, filters, blocks, max_batch=len(mine) * leaves, device=local_rank)
    all_boards = random_positions_device(n_total, seed=4242, min_plies=0, max_plies=40)     # every rank draws the same 512 openings, keeps its shard
    cfg = Config()
    games = [Game(cfg, board=all_boards[i]) for i in mine]
    rng = np.random.default_rng(11 + rank)
    K, W = args.steps, max(1, min(args.warmup, 2))

    def move_all(lv):
        res = cb_mcts.run_mcts_batch(games, net, sims, 30, True, None, rng=rng, leaves_per_tree=lv)
        c = net.engine.search_status()
        return res, c["evals"], c["sims"]

    # cost of virtual loss, untimed: the same searches without noise, sequential vs `leaves` per step — do they pick the same move?
    vl = None
    if leaves > 1 and not args.no_parity:
        class FirstTie:
            def choice(self, x):
                return x[0]
        sub = games[:min(64, len(games))]
        a = cb_mcts.run_mcts_batch(sub, net, sims, 0, False, None, rng=FirstTie(), leaves_per_tree=1)
        b = cb_mcts.run_mcts_batch(sub, net, sims, 0, False, None, rng=FirstTie(), leaves_per_tree=leaves)
        vl = {"positions": len(sub), "same_move_as_sequential": sum(x[0] == y[0] for x, y in zip(a, b)) / len(sub)}
    for _ in range(W):
        move_all(leaves)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    evals = nsims = 0
    for _ in range(K):
        res, e, s_ = move_all(leaves)
        evals += e; nsims += s_
        for g, (mv, _) in zip(games, res):
            if not g.terminal():
                g.make_move(mv)
        games[:] = [g if not g.terminal() else Game(cfg) for g in games]       # a finished game is replaced by a new one (train.py:54-79 loops)
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    (dt,), (evals, nsims) = _reduce(world, [dt], [evals, nsims])
    if rank == 0:
        emit(json.dumps({
            "metric": METRIC, "value": evals / dt, "unit": UNIT, "sims_per_s": nsims / dt, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"concurrent self-play (BASELINE config 3, literal): {n_total} games in total x {sims} sims per move, {blocks}x{filters} net, "
                                   f"sharded by game over {world} GPU(s) = {len(mine)} games per GPU, {leaves} leaf/leaves per tree per step "
                                   f"(evaluation batch {len(mine) * leaves}); a step = one move of every game",
                       "net": args.net, "games_total": n_total, "games_per_gpu": len(mine), "sims": sims, "leaves_per_tree": leaves,
                       "timing": "wall clock through mcts.run_mcts_batch: host objects in, moves played (end to end)"},
            "virtual_loss_cost": vl,
            "e2e": {"value": evals / dt, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "what": "the timed loop IS end to end"},
            "gpu_launches": K * ((sims + leaves - 1) // leaves + 1) * 4, "clocks": clocks, "cpu_baseline": None}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_train_arm(args):
    """BASELINE config 5: train.py step — 20x256 net, batch 4096 per GPU, data parallel with one NCCL
    all-reduce of the gradient buffer per step.  A step = forward + backward + all-reduce + SGD update on one synthetic batch."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from chessbot_b200.model import keras_init_weights
    from chessbot_b200.train import Trainer, train_flops_per_sample
    filters, blocks = parse_net(args.net)
    B, W, K = args.batch, args.warmup, args.steps
    tr = Trainer(keras_init_weights(filters, blocks, seed=0), filters, blocks, B, device=local_rank)
    rng = np.random.default_rng(100 + rank)
    images = rng.integers(0, 2, (B, 8, 8, 119), dtype=np.int16)            # synthetic planes of the sample format (train.py:104-110)
    images[..., 113] = rng.integers(0, 200, (B, 1, 1))
    images[..., 118] = rng.integers(0, 100, (B, 1, 1))
    z = rng.integers(-1, 2, B).astype(np.float32)
    pi = np.zeros((B, 4672), dtype=np.float32)
    cols = rng.integers(0, 4672, (B, 32))
    np.put_along_axis(pi, cols, rng.random((B, 32)).astype(np.float32), axis=1)
    pi /= pi.sum(1, keepdims=True)
    h_img, h_z, h_pi = (torch.from_numpy(a).pin_memory() for a in (images, z, pi))
    d_img, d_z, d_pi = (t.cuda() for t in (h_img, h_z, h_pi))
    lr = 0.01

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        tr.forward_backward(d_img, d_z, d_pi)
        tr.apply(lr)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):                                   # inputs resident in HBM
        tr.forward_backward(d_img, d_z, d_pi)
        tr.apply(lr)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    losses = tr.losses()
    tr.profile(True)
    for _ in range(max(2, K // 4)):
        tr.forward_backward(d_img, d_z, d_pi)
        tr.apply(lr)
    prof = tr.profile_read()
    tr.profile(False)
    prof_steps = max(2, K // 4)
    barrier()
    t0 = time.perf_counter()
    reps = max(2, K // 4)
    for _ in range(reps):                                # end to end: pinned host batch -> H2D -> step -> losses D2H
        out = tr.train_on_batch(h_img, h_z, h_pi, lr)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_dt = t.tolist()
    if rank == 0:
        peak_tf, peak_gbs, peak_src, peak_both = measured_peaks(ms * 1e-3)
        mma_ms = prof["conv_fwd"][0] + prof["wgrad"][0] + prof["dgrad"][0]
        tf = train_flops_per_sample(filters, blocks) * B * prof_steps / (mma_ms * 1e-3) / 1e12
        tot = sum(v[0] for v in prof.values())
        npad = (filters + 63) // 64 * 64
        act_mb = B / 2 * npad / 8 * 3200 / 1e6
        kernels = {k: {"launches": v[1], "ms_per_step": v[0] / prof_steps, "share_of_step": v[0] / tot} for k, v in prof.items() if v[1]}
        # HBM-bound BN passes, ALGORITHMIC bytes: one tensor pass = batch x 64 squares x padded channels x 2 B (the zero border
        # cells of the tile-native layout, +56 %, are not counted); passes per layer: forward = read Y, write A,
        # + read skip on the `blocks` residual layers; backward = (G, Y) twice + write dY, + (A twice, write dz) on those layers
        nl_ = 1 + 2 * blocks
        pass_bytes = B * 64 * npad * 2
        for k, passes in (("bn_fwd", 2 + blocks / nl_), ("bn_bwd", 5 + 3 * blocks / nl_)):
            gbs = passes * pass_bytes * nl_ / (prof[k][0] / prof_steps * 1e-3) / 1e9
            kernels[k].update({"bound": "hbm", "tensor_passes_per_layer": passes, "achieved_gbs": gbs, "peak_gbs": peak_gbs, "frac": gbs / peak_gbs})
        cpu = None if args.no_cpu_baseline else train_cpu_baseline_leg(filters, blocks)
        emit(json.dumps({
            "metric": "training samples/s (train.py step: forward + backward + gradient all-reduce + SGD update)", "value": world * B * K / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"train.py step (BASELINE config 5): {blocks}x{filters} net, batch {B} per GPU, BN batch statistics per GPU, "
                                   f"SGD nesterov + l2, gradients all-reduced over NCCL ({world} GPU(s))", "net": args.net, "batch_per_gpu": B,
                       "l2": "every layer's activation tensor (%.0f MB) exceeds the 126 MB L2; no flush" % act_mb},
            "roofline": {"bound": "tensor", "kernel": "tower_tcgen05_kernel (forward, data gradient) + wgrad_tcgen05_kernel", "achieved": tf,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "frac_of_burst_peak": tf / peak_both["burst"],
                         "frac_of_sustained_peak": tf / peak_both["sustained"], "traffic": None, "peak_source": peak_src,
                         "how": f"algorithmic FLOPs of the three 3x3-convolution passes / CUDA-event time of their launches over {prof_steps} steps",
                         "mma_share_of_step": mma_ms / tot},
            "kernels": kernels, "losses": losses,
            "e2e": {"value": world * B * reps / e2e_dt, "unit": "samples/s", "h2d_bytes_per_step": int(h_img.numel() * 2 + h_z.numel() * 4 + h_pi.numel() * 4),
                    "d2h_bytes_per_step": 24, "what": "Trainer.train_on_batch from pinned host arrays: H2D of images / z / pi, step, losses D2H"},
            "gpu_launches": int(sum(v[1] for v in prof.values()) / prof_steps) * K, "clocks": clocks, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


def run_reference_train_arm(args, filters, blocks, batch=16):
    """`--impl reference --workload train`: the reference's training step restated in torch fp32 (oracle/ref_train.py) on all host
    cores (torch intra-op threads), each step one bounded batch of `batch` samples."""
    import torch
    from oracle import ref_model, ref_train
    torch.set_num_threads(os.cpu_count() or 1)
    w = ref_model.init_weights(filters, blocks, seed=0)
    rng = np.random.default_rng(5)
    images = rng.integers(0, 2, (batch, 8, 8, 119)).astype(np.int16)
    z, pi = ref_train.synthetic_targets(batch, 5)
    vel = None
    for _ in range(max(args.warmup, 1)):
        _, w, vel, _ = ref_train.train_step(w, vel, images, z, pi, 0.01)
    steps = min(args.steps, 40)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, w, vel, _ = ref_train.train_step(w, vel, images, z, pi, 0.01)
    dt = time.perf_counter() - t0
    value = steps * batch / dt
    unit = "samples/s"
    sample = f"{steps} steps of batch {batch}, {blocks}x{filters} torch-fp32 autograd, {torch.get_num_threads()} threads (TensorFlow/Keras replaced by the oracle)"
    emit(json.dumps({
        "impl": "reference", "metric": "training samples/s (train.py step: forward + backward + gradient all-reduce + SGD update)", "value": value,
        "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"train.py step, {blocks}x{filters} net (reference CPU path, batch {batch})", "net": args.net},
        "cpu_baseline": {"value": value, "unit": unit, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def train_cpu_baseline_leg(filters, blocks, budget_s=15.0, batch=16):
    """The reference's training step restated in torch fp32 on the host cores (oracle/ref_train.py), bounded sample."""
    import torch
    from oracle import ref_model, ref_train
    w = ref_model.init_weights(filters, blocks, seed=0)
    rng = np.random.default_rng(5)
    images = rng.integers(0, 2, (batch, 8, 8, 119)).astype(np.int16)
    z, pi = ref_train.synthetic_targets(batch, 5)
    vel, steps = None, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s or steps < 1:
        _, w, vel, _ = ref_train.train_step(w, vel, images, z, pi, 0.01)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": steps * batch / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps of batch {batch}, {blocks}x{filters} torch-fp32 autograd on the host cores, {dt:.1f} s (TensorFlow/Keras replaced by the oracle)"}


_JSON_FD = None


def emit(line):
    """the ONE JSON line of the contract, written to the process's original stdout"""
    if _JSON_FD is None:
        print(line, flush=True)
    else:
        os.write(_JSON_FD, (line + "\n").encode())


def main():
    # NCCL prints its version banner (and anything at NCCL_DEBUG >= VERSION) on STDOUT; the bench prints exactly one JSON line
    # there.  So file descriptor 1 is pointed at stderr for everything else and the JSON line goes to the saved original.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--net", default="20x256", help="blocks x filters (BASELINE north_star: 20x256 at batch 1024)")
    ap.add_argument("--trees", type=int, default=1024, help="concurrent games per GPU = batch of leaves per step")
    ap.add_argument("--e2e-sims", type=int, default=100, help="simulations per search of the end-to-end leg (run_mcts default, mcts.py:38)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed checker leg (device network vs the torch-fp32 oracle)")
    ap.add_argument("--stream", default="hilo", help="residual stream between blocks: hilo = bf16 hi + lo (default), bf16 = round-1 datapath, "
                                                      "hilo@K = hi + lo only from residual block K on")
    ap.add_argument("--workload", default="search", choices=["search", "leaf_eval", "train", "puzzles", "selfplay"],
                    help="search = the headline (batched MCTS, BASELINE north_star); leaf_eval = BASELINE config 2 (use --net 6x128); "
                         "train = BASELINE config 5 (train.py step, --batch per GPU); puzzles = BASELINE config 4 (--positions x --sims); "
                         "selfplay = BASELINE config 3 read literally (--games in total x --sims, sharded over the GPUs)")
    ap.add_argument("--positions", type=int, default=10000, help="puzzles workload: positions in the sweep (all GPUs together)")
    ap.add_argument("--sims", type=int, default=None, help="puzzles / selfplay workloads: simulations per search (default 400 / 800)")
    ap.add_argument("--games", type=int, default=512, help="selfplay workload: concurrent games in total")
    ap.add_argument("--leaves", type=int, default=0, help="selfplay workload: leaves per tree per step (0 = 512 / games per GPU)")
    ap.add_argument("--agreement-positions", type=int, default=32, help="puzzles workload: positions of the untimed top-move agreement check")
    ap.add_argument("--batch", type=int, default=4096, help="train workload: samples per GPU per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "leaf_eval":
        run_leaf_eval_arm(args)
    elif args.workload == "train":
        run_train_arm(args)
    elif args.workload == "puzzles":
        args.sims = 400 if args.sims is None else args.sims
        run_puzzles_arm(args)
    elif args.workload == "selfplay":
        args.sims = 800 if args.sims is None else args.sims
        if "--steps" not in sys.argv:
            args.steps = 3
        run_selfplay_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
