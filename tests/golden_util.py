"""Loading of the committed golden fixtures (tests/golden/*.npz)."""
import json
import os

import numpy as np

import oracle_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAME_ID = {"othello": O.OG_OTHELLO, "c4": O.OG_C4, "go": O.OG_GO7, "go9": O.OG_GO9}
EVAL_ID = {"hash": O.OE_HASHNET, "hash1": O.OE_HASHNET, "uniform": O.OE_UNIFORM, "heuristic": O.OE_HEURISTIC}
HASH_SALT = {"hash1": 1}
INITQ_ID = {"parent": O.OQ_PARENT, "zero": O.OQ_ZERO, "drop": O.OQ_DROP_PARENT}

SELFPLAY_FIXTURES = sorted(f[len("selfplay_"):-4] for f in os.listdir(GOLDEN) if f.startswith("selfplay_"))
MATCH_FIXTURES = sorted(f[len("match_"):-4] for f in os.listdir(GOLDEN) if f.startswith("match_"))
ROLLOUT_FIXTURES = sorted(f[len("rollout_"):-4] for f in os.listdir(GOLDEN) if f.startswith("rollout_") and f.endswith(".npz"))
TREEWALK_FIXTURES = sorted(f[len("treewalk_"):-4] for f in os.listdir(GOLDEN) if f.startswith("treewalk_"))
GRIDNET_FIXTURES = sorted(f[len("gridnet_"):-4] for f in os.listdir(GOLDEN) if f.startswith("gridnet_"))
TREEWALK_KEYS = ["game_moves", "game_rng_draws", "move_N", "move_W", "move_P", "move_root_N", "move_root_W", "move_action",
                 "move_traversals", "move_player"]


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def rollout_digests(t, chunk):
    """SHA-256 per chunk of `chunk` games of every array of a rollout trace (same as tests/golden/make_golden.py)."""
    import hashlib
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


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    cmd = json.loads(str(d.pop("cmd")))
    if "states" in d:
        d["states"] = d["states"].astype(np.float32)
    return cmd, d


def match_agents(cmd):
    """The two agent descriptions of a match fixture, as oracle_py.match() takes them."""
    return [dict(evaluator=EVAL_ID[cmd["evaluators"][k]], hash_salt=HASH_SALT.get(cmd["evaluators"][k], 0),
                 use_sym=cmd["sym"][k], init_q=INITQ_ID[cmd["initq"][k]]) for k in range(2)]


def perft_table():
    with open(os.path.join(GOLDEN, "perft.json")) as f:
        return json.load(f)


def assert_trace_equal(ref, got, keys=None, exact=True):
    for k in (keys or ref.keys()):
        a, b = ref[k], got[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if exact:
            if not np.array_equal(a, b):
                bad = np.argwhere(a != b)
                raise AssertionError(f"{k}: {len(bad)} mismatches, first at {bad[0]}: ref {a[tuple(bad[0])]} got {b[tuple(bad[0])]}")
        else:
            np.testing.assert_allclose(b, a, rtol=0, atol=1e-6, err_msg=k)
