"""Throughput of the other BASELINE.json configurations on one B200 (bench.py measures config 3/4):
  config 2: Othello environment only -- perft(1..11) and random-rollout sweep (steps/s);
  config 1/5 shapes: Connect Four, Go 7x7 and Go 9x9 self-play with the reference's network shapes
  (library evaluator on tcgen05; SPRL_BENCH_EVALUATOR=libtorch runs the traced module through LibTorch instead).
Prints one JSON object per line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200 import capi, selfplay as SP
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network, trace_network

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)

# ---- config 1: Connect Four through the reference's own CPU worker, as is (C4Worker.cpp:11-27), one core, next to the GPU engine below
import subprocess
import tempfile
REF_WORKER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_worker")
if os.path.exists(REF_WORKER) and not os.environ.get("SPRL_BENCH_NO_CPU"):
    with tempfile.TemporaryDirectory() as tmp:
        pt = os.path.join(tmp, "c4.pt")
        trace_network(make_network("c4", 0), "cpu").save(pt)
        # iteration-0 leg: RandomNetwork, 2,048 descents, batch 1 / queue 1; iteration-1 leg: the traced net, 512 descents, 8 / 4
        for leg, model, sims, b, q in (("iteration0_uniform", "uniform", 2048, 1, 1), ("iteration1_traced_net", pt, 512, 8, 4)):
            out = subprocess.run([REF_WORKER, "c4", model, "0", "0", str(sims), str(b), str(q), "0.25", "0.5", "10", "10"],
                                 capture_output=True, text=True, check=True).stdout.strip().split("\n")[-1]
            o = json.loads(out)
            print(json.dumps({"config": "c4_reference_cpu_worker", "leg": leg, "cores": 1, "sims_per_move": sims, "batch_queue": [b, q],
                              "games": o["games"], "moves_per_sec": round(o["moves"] / o["seconds"], 2),
                              "sims_per_sec": round(o["sims"] / o["seconds"], 1), "window_seconds": round(o["seconds"], 1),
                              "protocol": "one discarded warm-up game, then up to 10 full games through the reference's runIteration "
                                          "(C4Worker.cpp constants) or 10 s, whichever ends first"}), flush=True)

# ---- config 2: perft
for d in (8, 9, 10, 11):
    count, ms = SP.env_perft(capi.GAME_OTHELLO, d)
    print(json.dumps({"config": "othello_perft", "depth": d, "leaves": count, "device_ms": round(ms, 3),
                      "leaves_per_sec": round(count / (ms / 1e3), 1)}), flush=True)
# ---- config 2: random rollouts, one thread per game
for lg in (10, 14, 18, 20, 22, 24):
    n = 1 << lg
    SP.env_rollout(capi.GAME_OTHELLO, 0, 0, min(n, 1 << 14))
    r = SP.env_rollout(capi.GAME_OTHELLO, 0, 0, n)
    steps = r["total_positions"] - n
    print(json.dumps({"config": "othello_rollout", "games": n, "steps": int(steps), "kernel_ms": round(r["elapsed_ms"], 3),
                      "steps_per_sec": round(steps / (r["elapsed_ms"] / 1e3), 1)}), flush=True)
for game, name in ((capi.GAME_C4, "c4"), (capi.GAME_GO7, "go7"), (capi.GAME_GO9, "go9")):
    r = SP.env_rollout(game, 0, 0, 1 << 18)
    steps = r["total_positions"] - (1 << 18)
    print(json.dumps({"config": name + "_rollout", "games": 1 << 18, "steps": int(steps), "kernel_ms": round(r["elapsed_ms"], 3),
                      "steps_per_sec": round(steps / (r["elapsed_ms"] / 1e3), 1)}), flush=True)

# ---- self-play of the other games (network through LibTorch on device buffers)
SCALE = int(os.environ.get("SPRL_BENCH_SLOT_SCALE", "1"))      # concurrent games = SCALE x the defaults below
for game, kind, sims, b, q, alpha, slots in ((capi.GAME_C4, "c4", 512, 8, 4, 0.5, 4096 * SCALE), (capi.GAME_GO7, "go7", 400, 16, 8, 0.2, 2048 * SCALE),
                                              (capi.GAME_GO9, "go9", 400, 16, 8, 0.2, 1024 * SCALE)):
    net = make_network(kind, 0)
    gi = capi.game_info(game)
    library = os.environ.get("SPRL_BENCH_EVALUATOR", "evalnet") != "libtorch"
    with SP.Engine(game, capi.EVAL_EXTERNAL, seed=0, sims=sims, max_batch=b, max_queue=q, dir_eps=0.25, dir_alpha=alpha,
                   num_slots=slots, max_games=slots * 8) as eng:
        if library:
            eng.attach_evalnet(EvalNet(net, device=0, rows=gi.rows, cols=gi.cols), use_cuda_graph=True)
        else:
            eng.attach_network(trace_network(net, dev), use_cuda_graph=True)
        eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        eng.begin_iteration(0, slots * 8)
        eng._capture()
        g = eng._nn["graph"]
        for _ in range(100):
            g.replay()
        torch.cuda.synchronize()
        eng.reset_stats()
        t0 = time.time()
        for _ in range(300):
            g.replay()
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = eng.stats()
        playing, failed = eng.poll()
    print(json.dumps({"config": kind + "_selfplay", "sims_per_move": sims, "batch_queue": [b, q], "slots": slots,
                      "sims_per_sec": round(st["sims"] / dt, 1), "moves_per_sec": round(st["moves"] / dt, 1),
                      "evals_per_sec": round(st["evals"] / dt, 1), "failed_slots": failed,
                      "evaluator": "library (tcgen05, fp16 split)" if library else "traced module, LibTorch/cuDNN fp32"}), flush=True)

# ---- config 3 throughput mode: Othello, 400 sims/move, leaf queue per tree swept at a constant leaf-batch capacity (65,536 rows)
for b, q in ((8, 4), (16, 8), (32, 16), (64, 32)):
    slots = 65536 // q
    with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=0, sims=400, max_batch=b, max_queue=q, dir_eps=0.25, dir_alpha=0.3,
                   num_slots=slots, max_games=slots * 8) as eng:
        eng.attach_evalnet(EvalNet(make_network("othello", 0), device=0), use_cuda_graph=True)
        eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        eng.begin_iteration(0, slots * 8)
        eng._capture()
        g = eng._nn["graph"]
        for _ in range(60):
            g.replay()
        torch.cuda.synchronize()
        eng.reset_stats()
        t0 = time.time()
        for _ in range(200):
            g.replay()
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = eng.stats()
        playing, failed = eng.poll()
    print(json.dumps({"config": "othello_queue_sweep", "sims_per_move": 400, "batch_queue": [b, q], "slots": slots,
                      "sims_per_sec": round(st["sims"] / dt, 1), "moves_per_sec": round(st["moves"] / dt, 1),
                      "evals_per_sim": round(st["evals"] / max(1, st["sims"]), 4), "rows_in_use": round(st["evals"] / 200 / 65536, 4),
                      "round_ms": round(dt / 200 * 1e3, 3), "failed_slots": failed}), flush=True)

# ---- match play (Evaluate.cpp): two random-init Othello networks, each side its own tree and evaluator
pairs, games = 4096, 8192
nets = [EvalNet(make_network("othello", k), device=0) for k in (0, 1)]
with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=0, sims=400, max_batch=8, max_queue=4, dir_eps=0.25, dir_alpha=0.1,
               u_weight=1.0, add_noise=1, num_slots=2 * pairs, max_games=games) as eng:
    eng.attach_match_evaluators(nets, use_cuda_graph=True)
    agents = [dict(evaluator=capi.EVAL_EXTERNAL, use_sym=1, init_q=capi.INITQ_PARENT)] * 2
    eng.reset_stats()
    torch.cuda.synchronize()
    t0 = time.time()
    res = eng.run_match(agents, games)
    torch.cuda.synchronize()
    dt = time.time() - t0
    st = eng.stats()
print(json.dumps({"config": "othello_match_play", "sims_per_move": 400, "batch_queue": [8, 4], "pairs_of_trees": pairs, "games": games,
                  "wins_agent0": res["wins"][0], "wins_agent1": res["wins"][1], "draws": res["draws"],
                  "games_per_sec": round(games / dt, 1), "sims_per_sec": round(st["sims"] / dt, 1),
                  "moves_per_sec": round(st["moves"] / dt, 1), "wall_s": round(dt, 2),
                  "evaluator": "two library evaluators (tcgen05, fp16 split), one per side"}), flush=True)
