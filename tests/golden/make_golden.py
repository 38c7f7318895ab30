"""Regenerate tests/golden/*.npz from the UNMODIFIED reference sources.

Run in the build container (where /root/reference exists):
    make -C oracle ref ref9 ref_gridnet && python tests/golden/make_golden.py
Each fixture is the output of oracle/_ref/ref_trace (reference sources compiled
verbatim + the contract RNG shim) for the command recorded in its `cmd` field.
Game "go9" = oracle/_ref/ref_trace9, the same tool over the build-dir copy of
games/GoNode.hpp with its two constants edited (9x9, komi 7.5; oracle/Makefile: ref9).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O  # noqa: E402

G_ID = {"othello": O.OG_OTHELLO, "c4": O.OG_C4, "go": O.OG_GO7, "go9": O.OG_GO9}

ROLLOUTS = [("othello", 1, 0, 24), ("c4", 1, 0, 48), ("go", 1, 0, 12), ("go9", 1, 0, 8)]
# BASELINE.json config 2: 2^16 Othello rollouts state by state -- too large to commit; the reference's arrays are
# committed as SHA-256 digests (rollout_othello_65536.json), chunk by chunk so that a mismatch is localised
ROLLOUT_DIGEST = ("othello", 0, 0, 1 << 16, 4096)
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
    # BASELINE.json config 5: Go 9x9 through the two-constant reference copy
    ("go9_hash_60_16_8", "go9", "hash", 0, 0, 1, 60, 16, 8, 0.25, 0.2, 1, 1, "parent"),
]
# The reference's LibTorch GridNetwork (its own glibc exp / mask / sequential sum / divide, networks/GridNetwork.hpp:107-142)
# over an integer-exact TorchScript module (tests/integer_net.py): name, game, net seed, seed, first, ngames, sims, b, q, eps, alpha
GRIDNET = [("othello_intnet_64_8_4", "othello", 3, 6, 0, 6, 64, 8, 4, 0.25, 0.3)]
# The step-wise surface: one tree per game through the public UCTTree API, the caller plays the first most-visited action
# name, game, evaluator, seed, first_game, ngames, sims, batch, queue, eps, alpha, noise, sym, initq
TREEWALK = [
    ("othello_hash_120_8_4", "othello", "hash", 3, 10, 3, 120, 8, 4, 0.25, 0.3, 1, 1, "parent"),
    ("c4_hash_200_8_4", "c4", "hash", 1, 0, 4, 200, 8, 4, 0.25, 0.5, 1, 1, "zero"),
    ("go_hash_60_16_8", "go", "hash", 2, 0, 1, 60, 16, 8, 0.25, 0.2, 1, 1, "parent"),
    ("go9_hash_40_16_8", "go9", "hash", 2, 0, 1, 40, 16, 8, 0.25, 0.2, 1, 1, "parent"),
]
NPY_SHAPES = [(5,), (3, 65), (2, 3, 8, 8), (16,), (1, 17, 7, 7)]
# match play (Evaluate.cpp).  name, game, evaluator0, evaluator1, seed, first_game, ngames, sims, batch,
# queue, sym0, initq0, sym1, initq1
MATCH = [
    ("othello_hash_hash1_100_8_4", "othello", "hash", "hash1", 0, 0, 2, 100, 8, 4, 1, "parent", 1, "parent"),
    ("othello_heuristic_uniform_60_8_4", "othello", "heuristic", "uniform", 1, 3, 2, 60, 8, 4, 0, "zero", 1, "parent"),
    ("c4_hash_hash1_128_8_4", "c4", "hash", "hash1", 0, 0, 4, 128, 8, 4, 1, "parent", 0, "zero"),
    ("go_hash_hash1_50_16_8", "go", "hash", "hash1", 0, 0, 1, 50, 16, 8, 1, "parent", 1, "drop"),
    ("go9_hash_hash1_40_16_8", "go9", "hash", "hash1", 0, 0, 1, 40, 16, 8, 1, "parent", 1, "zero"),
]
PERFT = {"othello": 9, "c4": 8, "go": 4, "go9": 3}


def tool_of(game):
    return ("ref_trace9", "go") if game == "go9" else ("ref_trace", game)


def digest_rollouts(t, chunk):
    """SHA-256 of every array of a rollout trace, per chunk of `chunk` games (positions are game-major)."""
    ends = np.cumsum(t["game_steps"].astype(np.int64))
    out = []
    for g0 in range(0, len(ends), chunk):
        lo = int(ends[g0 - 1]) if g0 else 0
        hi = int(ends[min(g0 + chunk, len(ends)) - 1])
        d = {k: hashlib.sha256(np.ascontiguousarray(t[k][lo:hi]).tobytes()).hexdigest() for k in ("cells", "player", "terminal", "winner", "mask", "action")}
        d["game_steps"] = hashlib.sha256(np.ascontiguousarray(t["game_steps"][g0:g0 + chunk]).tobytes()).hexdigest()
        d["positions"] = hi - lo
        out.append(d)
    return out


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    perft = {}
    for game, d in PERFT.items():
        tool, gname = tool_of(game)
        lines = O.run_ref("perft", gname, d, tool=tool).split("\n")
        perft[game] = [int(l.split()[1]) for l in lines if l.strip()]
    with open(os.path.join(HERE, "perft.json"), "w") as f:
        json.dump(perft, f, indent=1)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "t.trace")
        for game, seed, first, n in ROLLOUTS:
            tool, gname = tool_of(game)
            O.run_ref("rollout", gname, seed, first, n, path, tool=tool)
            t = O.read_trace(path)
            np.savez_compressed(os.path.join(HERE, f"rollout_{game}.npz"),
                                cmd=json.dumps(dict(game=game, seed=seed, first_game=first, ngames=n)), **t)
        for name, game, ev, seed, first, n, sims, b, q, eps, alpha, noise, sym, initq in SELFPLAY:
            tool, gname = tool_of(game)
            O.run_ref("selfplay", gname, ev, seed, first, n, sims, b, q, eps, alpha, noise, sym, initq, path, tool=tool)
            t = O.read_trace(path)
            # sample planes are 0/1: store as int8 to keep the fixture small
            t["states"] = t["states"].astype(np.int8)
            np.savez_compressed(os.path.join(HERE, f"selfplay_{name}.npz"),
                                cmd=json.dumps(dict(game=game, evaluator=ev, seed=seed, first_game=first, ngames=n,
                                                    sims=sims, max_batch=b, max_queue=q, eps=eps, alpha=alpha,
                                                    noise=noise, sym=sym, initq=initq)), **t)
        for name, game, ev0, ev1, seed, first, n, sims, b, q, sym0, q0, sym1, q1 in MATCH:
            tool, gname = tool_of(game)
            O.run_ref("match", gname, ev0, ev1, seed, first, n, sims, b, q, sym0, q0, sym1, q1, path, tool=tool)
            t = O.read_trace(path)
            np.savez_compressed(os.path.join(HERE, f"match_{name}.npz"),
                                cmd=json.dumps(dict(game=game, evaluators=[ev0, ev1], seed=seed, first_game=first, ngames=n,
                                                    sims=sims, max_batch=b, max_queue=q, sym=[sym0, sym1], initq=[q0, q1])), **t)
        for name, game, ev, seed, first, n, sims, b, q, eps, alpha, noise, sym, initq in TREEWALK:
            tool, gname = tool_of(game)
            O.run_ref("treewalk", gname, ev, seed, first, n, sims, b, q, eps, alpha, noise, sym, initq, path, tool=tool)
            t = O.read_trace(path)
            np.savez_compressed(os.path.join(HERE, f"treewalk_{name}.npz"),
                                cmd=json.dumps(dict(game=game, evaluator=ev, seed=seed, first_game=first, ngames=n, sims=sims,
                                                    max_batch=b, max_queue=q, eps=eps, alpha=alpha, noise=noise, sym=sym, initq=initq)), **t)
        # the 2^16-game rollout sweep of config 2, as digests
        game, seed, first, n, chunk = ROLLOUT_DIGEST
        O.run_ref("rollout", game, seed, first, n, path)
        t = O.read_trace(path)
        with open(os.path.join(HERE, f"rollout_{game}_{n}.json"), "w") as f:
            json.dump(dict(cmd=dict(game=game, seed=seed, first_game=first, ngames=n, chunk=chunk),
                           positions=int(t["game_steps"].sum()), chunks=digest_rollouts(t, chunk)), f, indent=1)
        # the reference's LibTorch GridNetwork over an integer-exact TorchScript module
        from integer_net import IntegerNet
        for name, game, net_seed, seed, first, n, sims, b, q, eps, alpha in GRIDNET:
            gi = O.game_info(G_ID[game])
            planes = 2 * gi.history + 1
            pt = os.path.join(tmp, "intnet.pt")
            IntegerNet(planes, gi.cells, gi.actions, seed=net_seed).save_torchscript(pt, planes, gi.rows, gi.cols)
            O.run_ref("selfplay", game, "pt:" + pt, seed, first, n, sims, b, q, eps, alpha, 1, 1, "parent", path, tool="ref_trace_torch")
            t = O.read_trace(path)
            t["states"] = t["states"].astype(np.int8)
            np.savez_compressed(os.path.join(HERE, f"gridnet_{name}.npz"),
                                cmd=json.dumps(dict(game=game, evaluator="intnet", net_seed=net_seed, seed=seed, first_game=first, ngames=n,
                                                    sims=sims, max_batch=b, max_queue=q, eps=eps, alpha=alpha, noise=1, sym=1, initq="parent")), **t)
        # the reference's own npy writer (utils/npy.hpp:616-639) on known arrays: the bytes, hex-encoded
        npy = {}
        for shape in NPY_SHAPES:
            O.run_ref("npy", "x", path, *shape)
            npy["x".join(map(str, shape))] = open(path, "rb").read().hex()
        with open(os.path.join(HERE, "npy_reference_bytes.json"), "w") as f:
            json.dump(npy, f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
