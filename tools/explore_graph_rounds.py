"""Exploration: rounds per CUDA graph.  Does capturing R search rounds (k_round, k_flip, k_evalnet, k_heads each) in ONE
graph instead of one round per graph shorten the round?  python tools/explore_graph_rounds.py [slots]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200 import capi, selfplay as SP
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network

dev = torch.device("cuda", 0)
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
net = make_network("othello", 0)
for R in (1, 4, 16, 1):
    with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=0, sims=400, max_batch=8, max_queue=4, dir_eps=0.25, dir_alpha=0.3,
                   num_slots=slots, max_games=slots * 16) as eng:
        ev = EvalNet(net, device=0)
        eng.attach_evalnet(ev, use_cuda_graph=True)
        eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        eng.begin_iteration(0, slots * 16)
        for _ in range(3):
            eng._forward()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
            for _ in range(R):
                eng._round_with_network()
        eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(768 // R):
            graph.replay()
        torch.cuda.synchronize()
        eng.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(768 // R):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = eng.stats()
        print(f"rounds/graph={R}: {ms / 768:.4f} ms per round, {st['sims'] / ms / 1e3:.2f} M sims/s", flush=True)
        # the same rounds as plain launches with events around each kernel group
        evs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(64)]
        for a, b, c in evs:
            a.record(); eng.round(); b.record(); eng._forward(); c.record()
        torch.cuda.synchronize()
        print(f"   probe: k_round {sum(a.elapsed_time(b) for a, b, c in evs) / 64:.4f} ms, forward {sum(b.elapsed_time(c) for a, b, c in evs) / 64:.4f} ms,"
              f" whole {evs[0][0].elapsed_time(evs[-1][2]) / 64:.4f} ms per round", flush=True)
        ev.close()
