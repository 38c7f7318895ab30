"""Small fixed workload of the kernels outside the search round, for ncu (config 2 and the sample writer):
perft (k_perft_expand / k_perft_count), random rollouts (k_rollout), batched transitions (k_env_step) and one
short self-play iteration with sample collection (k_begin / k_round / k_emit).

    python tools/profile_env.py [perft_depth] [log2_rollout_games] [selfplay_games]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sprl_b200 import capi, selfplay as SP

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 22
games = int(sys.argv[3]) if len(sys.argv) > 3 else 2048

count, ms = SP.env_perft(capi.GAME_OTHELLO, depth)
print(f"othello perft({depth}) = {count} in {ms:.3f} ms")
r = SP.env_rollout(capi.GAME_OTHELLO, 0, 0, 1 << lg)
steps = r["total_positions"] - (1 << lg)
print(f"othello rollouts: {1 << lg} games, {steps} steps in {r['elapsed_ms']:.3f} ms = {steps / r['elapsed_ms'] / 1e6:.2f} G steps/s")
for game, name in ((capi.GAME_C4, "c4"), (capi.GAME_GO7, "go7"), (capi.GAME_GO9, "go9")):
    r = SP.env_rollout(game, 0, 0, 1 << 18)
    print(f"{name} rollouts: {r['total_positions'] - (1 << 18)} steps in {r['elapsed_ms']:.3f} ms")
gi = capi.game_info(capi.GAME_OTHELLO)
n = 1 << 18
cells = np.full((n, gi.cells), -1, np.int8)
cells[:, 27] = cells[:, 36] = 1
cells[:, 28] = cells[:, 35] = 0
out = SP.env_step(capi.GAME_OTHELLO, cells, np.zeros(n, np.int8), np.full(n, 19, np.int32))
print("env_step:", n, "transitions, legal moves after the first move:", int(out["mask"][0].sum()))
with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_HASHNET, seed=0, sims=64, max_batch=8, max_queue=4, num_slots=games,
               max_games=games) as eng:
    s = eng.run_iteration(games)
    print("self-play iteration:", games, "games,", len(s[2]), "samples")
