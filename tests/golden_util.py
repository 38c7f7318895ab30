"""Loading of the committed golden fixtures (tests/golden/*.npz)."""
import json
import os

import numpy as np

import oracle_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAME_ID = {"othello": O.OG_OTHELLO, "c4": O.OG_C4, "go": O.OG_GO7}
EVAL_ID = {"hash": O.OE_HASHNET, "hash1": O.OE_HASHNET, "uniform": O.OE_UNIFORM, "heuristic": O.OE_HEURISTIC}
HASH_SALT = {"hash1": 1}
INITQ_ID = {"parent": O.OQ_PARENT, "zero": O.OQ_ZERO, "drop": O.OQ_DROP_PARENT}

SELFPLAY_FIXTURES = sorted(f[len("selfplay_"):-4] for f in os.listdir(GOLDEN) if f.startswith("selfplay_"))
MATCH_FIXTURES = sorted(f[len("match_"):-4] for f in os.listdir(GOLDEN) if f.startswith("match_"))
ROLLOUT_FIXTURES = sorted(f[len("rollout_"):-4] for f in os.listdir(GOLDEN) if f.startswith("rollout_"))


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
