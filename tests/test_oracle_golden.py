"""CPU tests: the plain-C oracle (oracle/liboracle.so) against golden vectors
produced by the UNMODIFIED reference sources (tests/golden/make_golden.py) and
against the reference's own known-answer test."""
import os

import numpy as np
import pytest

import golden_util as G
import oracle_py as O


def test_perft_matches_reference():
    table = G.perft_table()
    limits = {"othello": 8, "c4": 7, "go": 4, "go9": 3}      # keep the CPU suite short
    for game, counts in table.items():
        for d, want in enumerate(counts[:limits[game]], start=1):
            assert O.perft(G.GAME_ID[game], d) == want, (game, d)


def test_reference_c4_known_answer():
    # /root/reference/cpp/tests/test_c4.cpp:5-25: {3,3,4,4,2,3,1} -> player ZERO wins horizontally
    moves = [3, 3, 4, 4, 2, 3, 1]
    for k in range(len(moves)):
        r = O.replay(O.OG_C4, moves[:k])
        assert r["rc"] == 0 and r["terminal"] == 0 and r["winner"] == -1
        assert r["player"] == k % 2
    r = O.replay(O.OG_C4, moves)
    assert r["rc"] == 0 and r["terminal"] == 1
    assert r["rewards"].tolist() == [1.0, -1.0]
    assert r["mask"].sum() == 0        # terminal => no legal action (ConnectFourNode.cpp:70-72)


@pytest.mark.parametrize("name", G.ROLLOUT_FIXTURES)
def test_rollouts_match_reference(name):
    cmd, ref = G.load("rollout_" + name)
    got = O.rollout(G.GAME_ID[cmd["game"]], cmd["seed"], cmd["first_game"], cmd["ngames"])
    G.assert_trace_equal(ref, got)


@pytest.mark.parametrize("name", G.SELFPLAY_FIXTURES)
def test_selfplay_matches_reference(name):
    cmd, ref = G.load("selfplay_" + name)
    got = O.selfplay(G.GAME_ID[cmd["game"]], G.EVAL_ID[cmd["evaluator"]], cmd["seed"], cmd["first_game"],
                     cmd["ngames"], cmd["sims"], cmd["max_batch"], cmd["max_queue"], cmd["eps"], cmd["alpha"],
                     bool(cmd["noise"]), bool(cmd["sym"]), G.INITQ_ID[cmd["initq"]])
    # integer-valued and IEEE-exact quantities: bit-exact
    G.assert_trace_equal(ref, got, ["game_moves", "game_samples", "game_rng_draws", "move_N", "move_W", "move_P",
                                    "move_root_N", "move_root_W", "move_action", "move_traversals", "move_evals",
                                    "move_player", "move_board", "states", "outcomes"])
    # distributions go through pow(): the reference calls glibc powf, the contract a
    # deterministic double-based one -> 1e-6 absolute (observed: bit-identical)
    G.assert_trace_equal(ref, got, ["distributions"], exact=False)


@pytest.mark.parametrize("name", G.MATCH_FIXTURES)
def test_match_play_matches_reference(name):
    """Evaluate.cpp's match path (UCTNetworkAgent + playGame), two trees per game: bit-exact."""
    cmd, ref = G.load("match_" + name)
    got = O.match(G.GAME_ID[cmd["game"]], G.match_agents(cmd), cmd["seed"], cmd["first_game"], cmd["ngames"],
                  cmd["sims"], cmd["max_batch"], cmd["max_queue"])
    G.assert_trace_equal(ref, got, ["game_moves", "game_winner", "game_rng_draws", "move_N", "move_W", "move_P",
                                    "move_root_N", "move_root_W", "move_action", "move_traversals", "move_agent",
                                    "move_player"])
    first = ref["game_first"]
    wins = [int(((ref["game_winner"] >= 0) & ((ref["game_winner"] ^ first) == k)).sum()) for k in (0, 1)]
    assert got["wins"] == tuple(wins) and got["draws"] == int((ref["game_winner"] < 0).sum())


@pytest.mark.parametrize("name", G.TREEWALK_FIXTURES)
def test_treewalk_matches_reference(name):
    """One tree per game through the reference's public UCTTree API, the caller playing the first most-visited action
    (ref_trace `treewalk`): what the step-wise C ABI (sprl_search / sprl_root_stats / sprl_advance) is checked against."""
    cmd, ref = G.load("treewalk_" + name)
    got = O.selfplay(G.GAME_ID[cmd["game"]], G.EVAL_ID[cmd["evaluator"]], cmd["seed"], cmd["first_game"], cmd["ngames"],
                     cmd["sims"], cmd["max_batch"], cmd["max_queue"], cmd["eps"], cmd["alpha"], bool(cmd["noise"]),
                     bool(cmd["sym"]), G.INITQ_ID[cmd["initq"]], caller_moves=True)
    G.assert_trace_equal(ref, got, G.TREEWALK_KEYS)


@pytest.mark.parametrize("name", G.GRIDNET_FIXTURES)
def test_gridnetwork_postprocess_matches_reference(name):
    """The fixture went through the verbatim reference's LibTorch GridNetwork (networks/GridNetwork.hpp:62-145: embed,
    forward, glibc exp, mask, sequential sum, multiply by the reciprocal) over an integer-exact TorchScript module; the
    oracle evaluates the same function through its callback and the contract's deterministic exp.  Root priors within
    1e-6 (the two exps may differ in the last bit), and with these inputs the whole search agrees bit for bit."""
    from integer_net import IntegerNet
    cmd, ref = G.load("gridnet_" + name)
    gi = O.game_info(G.GAME_ID[cmd["game"]])
    net = IntegerNet(2 * gi.history + 1, gi.cells, gi.actions, seed=cmd["net_seed"])
    got = O.selfplay(G.GAME_ID[cmd["game"]], O.OE_CALLBACK, cmd["seed"], cmd["first_game"], cmd["ngames"], cmd["sims"],
                     cmd["max_batch"], cmd["max_queue"], cmd["eps"], cmd["alpha"], eval_fn=net.numpy)
    first = np.concatenate([[0], np.cumsum(ref["game_moves"])[:-1]])
    np.testing.assert_allclose(got["move_P"][first], ref["move_P"][first], rtol=0, atol=1e-6)
    G.assert_trace_equal(ref, got, ["game_moves", "game_rng_draws", "move_N", "move_action", "move_traversals", "move_player", "states", "outcomes"])
    np.testing.assert_allclose(got["move_P"], ref["move_P"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["move_W"], ref["move_W"], rtol=0, atol=1e-5)


def test_rollout_sweep_digests_match_reference():
    """BASELINE.json config 2: the first 4,096 of the 2^16 Othello rollouts, state by state, as SHA-256 digests of the
    verbatim reference's arrays (the GPU suite checks all 16 chunks)."""
    fx = G.load_json("rollout_othello_65536.json")
    c = fx["cmd"]
    got = O.rollout(G.GAME_ID[c["game"]], c["seed"], c["first_game"], c["chunk"])
    assert G.rollout_digests(got, c["chunk"])[0] == fx["chunks"][0]


def test_npy_writer_bytes_match_reference():
    """sprl_write_npy_f32 (the C ABI's writer; host code, no GPU needed) and the oracle's writer against the bytes the
    reference's own npy::write_npy (utils/npy.hpp:616-639) produced for the same arrays."""
    import ctypes as C
    import tempfile
    from sprl_b200 import capi
    fx = G.load_json("npy_reference_bytes.json")
    assert len(fx) >= 5
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "a.npy")
        for key, hexbytes in fx.items():
            shape = tuple(int(x) for x in key.split("x"))
            a = (np.arange(int(np.prod(shape)), dtype=np.float32) * np.float32(0.25) - np.float32(3.0)).reshape(shape)
            want = bytes.fromhex(hexbytes)
            dims = (C.c_uint64 * len(shape))(*shape)
            capi.check(capi.load().sprl_write_npy_f32(path.encode(), a.ctypes.data_as(C.c_void_p), dims, len(shape)))
            assert open(path, "rb").read() == want, key
            O.write_npy(path, a)
            assert open(path, "rb").read() == want, key
            assert np.array_equal(np.load(path), a)


def test_othello_quirks():
    """Quirks the survey lists are visible in the traces."""
    cmd, ref = G.load("selfplay_othello_hash_400_8_4")
    # Q7: whole batches are added, so >= 400 and < 408 traversals per move
    assert (ref["move_traversals"] >= 400).all() and (ref["move_traversals"] < 408).all()
    # Q4: the fresh root is queued max_queue times in the first batch of a game:
    # root N = traversals but children sum to traversals - 4
    assert ref["move_root_N"][0] == ref["move_traversals"][0]
    assert ref["move_N"][0].sum() == ref["move_traversals"][0] - 4
    # Q2: later roots keep the visit count of the edge that led to them
    assert ref["move_root_N"][1] > ref["move_traversals"][1]


def test_npy_writer_header(tmp_path):
    # utils/npy.hpp:430-476: 16-byte alignment, "(N,)" for 1-D, v1.0
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2)
    p = str(tmp_path / "a.npy")
    O.write_npy(p, a)
    assert np.array_equal(np.load(p), a)
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x93NUMPY\x01\x00"
    hlen = raw[8] + 256 * raw[9]
    assert (10 + hlen) % 16 == 0 and raw[10 + hlen - 1:10 + hlen] == b"\n"
    assert raw[10:10 + hlen].startswith(b"{'descr': '<f4', 'fortran_order': False, 'shape': (2, 3, 2, 2), }")
    b = np.arange(5, dtype=np.float32)
    O.write_npy(p, b)
    assert b"'shape': (5,), }" in open(p, "rb").read()[:80]
    assert np.array_equal(np.load(p), b)


def test_contract_math():
    # deterministic pow/exp stay within 1 ulp of libm over the ranges the path uses
    xs = np.linspace(0.0, 1.0, 2001, dtype=np.float32)
    for e in (0.98, 10.0):
        got = np.array([O.lib().oracle_det_powf(float(x), e) for x in xs], np.float32)
        want = np.power(xs.astype(np.float64), np.float64(np.float32(e))).astype(np.float32)
        assert np.all(np.abs(got.view(np.int32) - want.view(np.int32)) <= 1)
    ls = np.linspace(-30, 30, 1201, dtype=np.float32)
    got = np.array([O.lib().oracle_det_expf(float(x)) for x in ls], np.float32)
    want = np.exp(ls.astype(np.float64)).astype(np.float32)
    assert np.all(np.abs(got.view(np.int32) - want.view(np.int32)) <= 1)


def test_dirichlet_statistics():
    # Dirichlet(alpha) over n slots: mean 1/n, var (n-1)/(n^2 (n alpha + 1))
    import ctypes as C
    n, alpha, reps = 8, 0.3, 4000
    out = np.zeros(n, np.float32)
    acc = np.zeros((reps, n))
    for g in range(reps):
        O.lib().oracle_dirichlet(C.c_uint64(11), C.c_uint64(g), C.c_uint64(0), C.c_float(alpha),
                                 out.ctypes.data_as(C.c_void_p), C.c_int(n), None)
        acc[g] = out
    assert np.allclose(acc.sum(1), 1.0, atol=1e-5)
    assert np.allclose(acc.mean(0), 1.0 / n, atol=0.012)
    var = (n - 1) / (n * n * (n * alpha + 1))
    assert np.allclose(acc.var(0), var, rtol=0.15)
