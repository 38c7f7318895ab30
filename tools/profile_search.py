"""Small fixed workload for ncu: a fixed number of search launches over G concurrent Othello games.

    python tools/profile_search.py G launches {uniform|hash|ext} rounds_per_launch [warm_rounds]

`ext` is the benchmark's configuration: SPRL_EVAL_EXTERNAL with the 2x64 network run by the library's
evaluator between search launches (no CUDA graph, so every launch is a plain kernel for ncu);
`warm_rounds` search rounds are played first so the profiled launches see mid-game trees."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprl_b200 import capi, selfplay as SP

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kind = sys.argv[3] if len(sys.argv) > 3 else "uniform"
rpl = int(sys.argv[4]) if len(sys.argv) > 4 else 8
warm = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ev = {"uniform": capi.EVAL_UNIFORM, "hash": capi.EVAL_HASHNET, "ext": capi.EVAL_EXTERNAL}[kind]
with SP.Engine(capi.GAME_OTHELLO, ev, sims=400, max_batch=8, max_queue=4, num_slots=G, max_games=G * 4,
               rounds_per_launch=rpl) as eng:
    if kind == "ext":
        import torch
        from sprl_b200.evalnet import EvalNet
        from sprl_b200.network import make_network
        eng.attach_evalnet(EvalNet(make_network("othello", 0), device=0), use_cuda_graph=False)
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        step = eng._round_with_network
    else:
        step = eng.round
    eng.begin_iteration(0, G * 4)
    for _ in range(warm):
        step()
    eng.poll()
    eng.reset_stats()
    t = time.time()
    for _ in range(launches):
        step()
    playing, failed = eng.poll()
    dt = time.time() - t
    st = eng.stats()
    print(f"G={G} launches={launches} kind={kind} rpl={rpl} warm={warm}: {dt*1e3:.1f} ms, {st['sims']/dt/1e6:.2f} M sims/s, "
          f"moves {st['moves']}, playing {playing}, failed {failed}")
