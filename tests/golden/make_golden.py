"""Regenerate tests/golden/*.npz from the UNMODIFIED reference sources.

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py
Each fixture is the output of oracle/_ref/ref_trace (reference sources compiled
verbatim + the contract RNG shim) for the command recorded in its `cmd` field.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O  # noqa: E402

ROLLOUTS = [("othello", 1, 0, 24), ("c4", 1, 0, 48), ("go", 1, 0, 12)]
# name, game, evaluator, seed, first_game, ngames, sims, batch, queue, eps, alpha, noise, sym, initq
SELFPLAY = [
    ("othello_hash_400_8_4", "othello", "hash", 0, 0, 2, 400, 8, 4, 0.25, 0.3, 1, 1, "parent"),
    ("othello_uniform_64_1_1", "othello", "uniform", 5, 3, 1, 64, 1, 1, 0.25, 0.3, 1, 1, "parent"),
    ("othello_hash_plain_100", "othello", "hash", 9, 7, 1, 100, 8, 4, 0.25, 0.3, 0, 0, "zero"),
    ("othello_hash_200_16_16", "othello", "hash", 2, 40, 1, 200, 16, 16, 0.25, 0.3, 1, 1, "parent"),
    ("c4_hash_512_8_4", "c4", "hash", 0, 0, 3, 512, 8, 4, 0.25, 0.5, 1, 1, "parent"),
    ("c4_uniform_2048_1_1", "c4", "uniform", 1, 0, 1, 2048, 1, 1, 0.25, 0.5, 1, 1, "parent"),
    ("go_hash_100_16_8", "go", "hash", 0, 0, 1, 100, 16, 8, 0.25, 0.2, 1, 1, "parent"),
    # networks/OthelloHeuristic.cpp as the evaluator, InitQ::DROP_PARENT
    ("othello_heuristic_100_8_4", "othello", "heuristic", 3, 0, 1, 100, 8, 4, 0.25, 0.3, 1, 1, "parent"),
    ("othello_hash_drop_120_8_4", "othello", "hash", 4, 0, 1, 120, 8, 4, 0.25, 0.3, 1, 1, "drop"),
    ("c4_hash_drop_200_8_4", "c4", "hash", 2, 5, 2, 200, 8, 4, 0.25, 0.5, 1, 1, "drop"),
]
# match play (Evaluate.cpp).  name, game, evaluator0, evaluator1, seed, first_game, ngames, sims, batch,
# queue, sym0, initq0, sym1, initq1
MATCH = [
    ("othello_hash_hash1_100_8_4", "othello", "hash", "hash1", 0, 0, 2, 100, 8, 4, 1, "parent", 1, "parent"),
    ("othello_heuristic_uniform_60_8_4", "othello", "heuristic", "uniform", 1, 3, 2, 60, 8, 4, 0, "zero", 1, "parent"),
    ("c4_hash_hash1_128_8_4", "c4", "hash", "hash1", 0, 0, 4, 128, 8, 4, 1, "parent", 0, "zero"),
    ("go_hash_hash1_50_16_8", "go", "hash", "hash1", 0, 0, 1, 50, 16, 8, 1, "parent", 1, "drop"),
]
PERFT = {"othello": 9, "c4": 8, "go": 3}


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    perft = {}
    for game, d in PERFT.items():
        lines = O.run_ref("perft", game, d).split("\n")
        perft[game] = [int(l.split()[1]) for l in lines if l.strip()]
    with open(os.path.join(HERE, "perft.json"), "w") as f:
        json.dump(perft, f, indent=1)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "t.trace")
        for game, seed, first, n in ROLLOUTS:
            O.run_ref("rollout", game, seed, first, n, path)
            t = O.read_trace(path)
            np.savez_compressed(os.path.join(HERE, f"rollout_{game}.npz"),
                                cmd=json.dumps(dict(game=game, seed=seed, first_game=first, ngames=n)), **t)
        for name, game, ev, seed, first, n, sims, b, q, eps, alpha, noise, sym, initq in SELFPLAY:
            O.run_ref("selfplay", game, ev, seed, first, n, sims, b, q, eps, alpha, noise, sym, initq, path)
            t = O.read_trace(path)
            # sample planes are 0/1: store as int8 to keep the fixture small
            t["states"] = t["states"].astype(np.int8)
            np.savez_compressed(os.path.join(HERE, f"selfplay_{name}.npz"),
                                cmd=json.dumps(dict(game=game, evaluator=ev, seed=seed, first_game=first, ngames=n,
                                                    sims=sims, max_batch=b, max_queue=q, eps=eps, alpha=alpha,
                                                    noise=noise, sym=sym, initq=initq)), **t)
        for name, game, ev0, ev1, seed, first, n, sims, b, q, sym0, q0, sym1, q1 in MATCH:
            O.run_ref("match", game, ev0, ev1, seed, first, n, sims, b, q, sym0, q0, sym1, q1, path)
            t = O.read_trace(path)
            np.savez_compressed(os.path.join(HERE, f"match_{name}.npz"),
                                cmd=json.dumps(dict(game=game, evaluators=[ev0, ev1], seed=seed, first_game=first, ngames=n,
                                                    sims=sims, max_batch=b, max_queue=q, sym=[sym0, sym1], initq=[q0, q1])), **t)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
