"""The workload for compute-sanitizer (memcheck / racecheck are this engine's analogue of the reference's
absent race detection): 64 Othello trees through k_round -> k_flip -> k_evalnet_resident x2 -> k_heads for a few dozen
rounds (no CUDA graph), then the sample writer.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py [rounds]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprl_b200 import capi
from sprl_b200 import selfplay as SP
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ev = EvalNet(make_network("othello", 0), device=0)
if len(sys.argv) > 2:
    ev.set_path(int(sys.argv[2]))
with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=1, sims=24, max_batch=8, max_queue=4, num_slots=64, max_games=64) as eng:
    eng.attach_evalnet(ev, use_cuda_graph=False)
    import torch
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.begin_iteration(0, 64)
    for _ in range(rounds):
        eng._round_with_network()
    playing, failed = eng.poll()
    st = eng.stats()
    ev.status()
    print("rounds", rounds, "playing", playing, "failed", failed, "sims", st["sims"], "moves", st["moves"], "phases", ev.phases, flush=True)
    assert failed == 0 and st["moves"] > 0
