"""Exploration: does splitting the slots of one GPU into G engine groups on G streams (search of one
group overlapping the evaluator of another) raise throughput?  python tools/explore_overlap.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200 import capi, selfplay as SP
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network

dev = torch.device("cuda", 0)
net = make_network("othello", 0)
TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 8192

GROUPS = [int(g) for g in os.environ.get("GROUPS", "1,2,3,4").split(",")]
print("SPRL_EVALNET_MAX_CTAS =", os.environ.get("SPRL_EVALNET_MAX_CTAS"), flush=True)
for groups in GROUPS:
    slots = TOTAL // groups
    engines, graphs, streams, evs = [], [], [], []
    for g in range(groups):
        ev = EvalNet(net, device=0)
        eng = SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=0, sims=400, max_batch=8, max_queue=4, num_slots=slots, max_games=slots * 24)
        eng.set_game_stride(groups)
        eng.attach_evalnet(ev, use_cuda_graph=True)
        s = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(s):
            eng.set_stream(s.cuda_stream)
            eng.begin_iteration(g, slots * 24)
            eng._capture()
            eng.set_stream(s.cuda_stream)
        engines.append(eng); graphs.append(eng._nn["graph"]); streams.append(s); evs.append(ev)
    torch.cuda.synchronize()

    def rounds(n):
        for _ in range(n):
            for g in range(groups):
                with torch.cuda.stream(streams[g]):
                    graphs[g].replay()

    rounds(600)
    torch.cuda.synchronize()
    for e in engines:
        e.reset_stats()
    t0 = time.time()
    rounds(400)
    torch.cuda.synchronize()
    dt = time.time() - t0
    sims = sum(e.stats()["sims"] for e in engines)
    moves = sum(e.stats()["moves"] for e in engines)
    print(f"groups={groups} slots/group={slots}: {sims / dt / 1e6:.2f} M sims/s, {moves / dt:.0f} moves/s, {dt / 400 * 1e3:.3f} ms per round of all groups", flush=True)
    for e in engines:
        e.close()
    for ev in evs:
        ev.close()
