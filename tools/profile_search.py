"""Small fixed workload for ncu: search-only self-play (device evaluator), a fixed number of launches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprl_b200 import capi, selfplay as SP
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ev = capi.EVAL_HASHNET if (len(sys.argv) > 3 and sys.argv[3] == "hash") else capi.EVAL_UNIFORM
rpl = int(sys.argv[4]) if len(sys.argv) > 4 else 8
with SP.Engine(capi.GAME_OTHELLO, ev, sims=400, max_batch=8, max_queue=4, num_slots=G, max_games=G, rounds_per_launch=rpl) as eng:
    eng.begin_iteration(0, G)
    t = time.time()
    for _ in range(launches):
        eng.round()
    playing, failed = eng.poll()
    dt = time.time() - t
    st = eng.stats()
    print(f"G={G} launches={launches} rpl={rpl}: {dt*1e3:.1f} ms, {st['sims']/dt/1e6:.1f} M sims/s, moves {st['moves']}, playing {playing}, failed {failed}")
