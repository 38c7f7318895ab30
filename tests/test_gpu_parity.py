"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs and against the golden fixtures
produced by the unmodified reference.  Integer / index work is compared bit-exact;
fp32 search statistics are compared bit-exact too (the kernels keep the reference's
operand order, no FMA); sampled distributions within 1e-6 (pow)."""
import numpy as np
import pytest

import golden_util as G
import oracle_py as O
from sprl_b200 import capi
from sprl_b200 import selfplay as SP

pytestmark = pytest.mark.gpu

GAMES = {"othello": capi.GAME_OTHELLO, "c4": capi.GAME_C4, "go": capi.GAME_GO7}
EVALS = {"hash": capi.EVAL_HASHNET, "uniform": capi.EVAL_UNIFORM}
INITQ = {"parent": capi.INITQ_PARENT, "zero": capi.INITQ_ZERO}

EXACT_KEYS = ["game_moves", "game_rng_draws", "move_N", "move_W", "move_P", "move_root_N", "move_root_W",
              "move_action", "move_traversals", "move_player", "states", "outcomes"]


def test_device_present():
    assert capi.load().sprl_device_count() > 0, "no CUDA device: the GPU tests must not pass on a fallback"


# ------------------------------------------------------------------ environment
def test_perft_bit_exact():
    table = G.perft_table()
    for name, counts in table.items():
        for d, want in enumerate(counts, start=1):
            got, _ = SP.env_perft(GAMES[name], d)
            assert got == want, (name, d, got, want)
    # Othello perft(10): confirmed against the verbatim reference during the survey
    assert SP.env_perft(capi.GAME_OTHELLO, 10)[0] == 24571284
    # Othello perft(11): published value
    assert SP.env_perft(capi.GAME_OTHELLO, 11)[0] == 212258800
    # Go 9x9: against the oracle's two-constant variant
    for d in (1, 2, 3):
        assert SP.env_perft(capi.GAME_GO9, d)[0] == O.perft(O.OG_GO9, d)
    assert SP.env_perft(capi.GAME_OTHELLO, 0)[0] == 1


@pytest.mark.parametrize("name", G.ROLLOUT_FIXTURES)
def test_rollout_golden(name):
    cmd, ref = G.load("rollout_" + name)
    got = SP.env_rollout(GAMES[cmd["game"]], cmd["seed"], cmd["first_game"], cmd["ngames"], record=True)
    G.assert_trace_equal(ref, got)


@pytest.mark.parametrize("game,n", [(capi.GAME_OTHELLO, 2000), (capi.GAME_C4, 3000), (capi.GAME_GO7, 500), (capi.GAME_GO9, 120)])
def test_rollout_vs_oracle_state_by_state(game, n):
    got = SP.env_rollout(game, 77, 1000, n, record=True)
    ref = O.rollout(game, 77, 1000, n)
    G.assert_trace_equal(ref, got)
    assert got["total_positions"] == len(ref["player"])


def test_rollout_large_sweep_properties():
    n = 1 << 18
    r = SP.env_rollout(capi.GAME_OTHELLO, 0, 0, n)
    assert r["game_steps"].min() >= 10 and r["game_steps"].max() <= 129
    assert r["total_positions"] == int(r["game_steps"].astype(np.int64).sum())
    # the first games of the sweep are the games of a small, recorded run
    small = SP.env_rollout(capi.GAME_OTHELLO, 0, 0, 64, record=True)
    assert np.array_equal(small["game_steps"], r["game_steps"][:64])
    assert np.array_equal(small["final_winner"], r["final_winner"][:64])
    assert set(np.unique(r["final_winner"])) <= {-1, 0, 1}


def test_env_step_matches_oracle_positions():
    for game in (O.OG_OTHELLO, O.OG_C4):
        ref = O.rollout(game, 3, 0, 200)
        idx = np.nonzero(ref["action"] >= 0)[0]
        out = SP.env_step(game, ref["cells"][idx], ref["player"][idx], ref["action"][idx])
        nxt = idx + 1
        assert np.array_equal(out["cells"], ref["cells"][nxt])
        assert np.array_equal(out["player"], ref["player"][nxt])
        assert np.array_equal(out["terminal"], ref["terminal"][nxt])
        assert np.array_equal(out["winner"], ref["winner"][nxt])
        assert np.array_equal(out["mask"], ref["mask"][nxt])
    assert SP.env_step(capi.GAME_OTHELLO, np.zeros((0, 64), np.int8), np.zeros(0, np.int8), np.zeros(0, np.int32))["cells"].shape == (0, 64)


# -------------------------------------------------------------------- self-play
def run_engine(game, evaluator, seed, first_game, ngames, sims, b, q, eps, alpha, noise, sym, initq, num_slots=None, **kw):
    with SP.Engine(game, evaluator, seed=seed, sims=sims, max_batch=b, max_queue=q, dir_eps=eps, dir_alpha=alpha,
                   add_noise=int(noise), use_sym=int(sym), init_q=initq, num_slots=num_slots or ngames,
                   max_games=ngames, record_stats=1, **kw) as eng:
        states, dists, outcomes = eng.run_iteration(ngames, first_game=first_game)
        got = eng.move_stats(ngames)
        got.update(states=states, distributions=dists, outcomes=outcomes, stats=eng.stats())
    return got


def compare_selfplay(ref, got):
    G.assert_trace_equal(ref, got, EXACT_KEYS)
    G.assert_trace_equal(ref, got, ["distributions"], exact=False)


@pytest.mark.parametrize("name", G.SELFPLAY_FIXTURES)
def test_selfplay_golden(name):
    cmd, ref = G.load("selfplay_" + name)
    got = run_engine(GAMES[cmd["game"]], EVALS[cmd["evaluator"]], cmd["seed"], cmd["first_game"], cmd["ngames"],
                     cmd["sims"], cmd["max_batch"], cmd["max_queue"], cmd["eps"], cmd["alpha"], cmd["noise"],
                     cmd["sym"], INITQ[cmd["initq"]])
    compare_selfplay(ref, got)


@pytest.mark.parametrize("game,ev,sims,b,q,alpha,ngames,slots", [
    (capi.GAME_OTHELLO, "hash", 400, 8, 4, 0.3, 24, 24),
    (capi.GAME_OTHELLO, "hash", 100, 8, 4, 0.3, 40, 7),       # slots reused by several games
    (capi.GAME_OTHELLO, "uniform", 64, 1, 1, 0.3, 16, 16),
    (capi.GAME_OTHELLO, "hash", 150, 32, 32, 0.3, 8, 8),      # wide queue
    (capi.GAME_C4, "hash", 512, 8, 4, 0.5, 24, 5),
    (capi.GAME_C4, "uniform", 200, 8, 4, 0.5, 16, 16),
    (capi.GAME_GO7, "hash", 100, 16, 8, 0.2, 6, 6),
    (capi.GAME_GO9, "hash", 60, 16, 8, 0.2, 3, 3),
])
def test_selfplay_vs_oracle(game, ev, sims, b, q, alpha, ngames, slots):
    seed, first = 1234, 500
    ref = O.selfplay(game, O.OE_HASHNET if ev == "hash" else O.OE_UNIFORM, seed, first, ngames, sims, b, q, 0.25, alpha,
                     max_moves_per_game=170)
    got = run_engine(game, EVALS[ev], seed, first, ngames, sims, b, q, 0.25, alpha, 1, 1, capi.INITQ_PARENT, num_slots=slots)
    compare_selfplay(ref, got)
    st, rs = got["stats"], ref["stats"]
    assert st["sims"] == rs["total_traversals"] and st["evals"] == rs["total_evals"]
    assert st["leaves_terminal"] == rs["leaves_terminal"] and st["leaves_gray"] == rs["leaves_gray"]
    assert st["depth_sum"] == int(rs["select_depth_sum"]) and st["legal_sum"] == int(rs["select_legal_sum"])


def test_selfplay_variants_vs_oracle():
    # no symmetrizer, no noise, InitQ::ZERO
    for noise, sym, initq, oq in [(0, 0, capi.INITQ_ZERO, O.OQ_ZERO), (1, 0, capi.INITQ_PARENT, O.OQ_PARENT), (0, 1, capi.INITQ_ZERO, O.OQ_ZERO)]:
        ref = O.selfplay(O.OG_OTHELLO, O.OE_HASHNET, 9, 0, 4, 80, 8, 4, 0.25, 0.3, add_noise=bool(noise), use_sym=bool(sym), init_q=oq)
        got = run_engine(capi.GAME_OTHELLO, capi.EVAL_HASHNET, 9, 0, 4, 80, 8, 4, 0.25, 0.3, noise, sym, initq)
        compare_selfplay(ref, got)


def test_results_do_not_depend_on_slot_count():
    a = run_engine(capi.GAME_OTHELLO, capi.EVAL_HASHNET, 5, 0, 12, 60, 8, 4, 0.25, 0.3, 1, 1, capi.INITQ_PARENT, num_slots=12)
    b = run_engine(capi.GAME_OTHELLO, capi.EVAL_HASHNET, 5, 0, 12, 60, 8, 4, 0.25, 0.3, 1, 1, capi.INITQ_PARENT, num_slots=5,
                   rounds_per_launch=3)
    for k in EXACT_KEYS + ["distributions"]:
        assert np.array_equal(a[k], b[k]), k


def test_capacity_overflow_is_reported():
    with pytest.raises(capi.SprlError) as ei:
        run_engine(capi.GAME_OTHELLO, capi.EVAL_HASHNET, 5, 0, 2, 400, 8, 4, 0.25, 0.3, 1, 1, capi.INITQ_PARENT, units_per_tree=600)
    assert ei.value.code == capi.SPRL_E_CAPACITY


def test_sample_round_trip_properties():
    """Size-independent properties at a larger scale: every sample block of 8 is the D4 orbit
    of its first element; distributions sum to 1; outcomes are in {-1,0,1} and flip with the mover."""
    with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_UNIFORM, seed=3, sims=50, max_batch=8, max_queue=4, num_slots=256, max_games=512) as eng:
        states, dists, outcomes = eng.run_iteration(512)
        st = eng.stats()
    assert st["games"] == 512 and states.shape[0] == 8 * st["moves"]
    assert np.allclose(dists.sum(1), 1.0, atol=1e-5)
    assert set(np.unique(outcomes)) <= {-1.0, 0.0, 1.0}
    s = states.reshape(-1, 8, 3, 8, 8)
    assert np.array_equal(s[:, 4], s[:, 0][..., ::-1])                      # symmetry 4 = column flip
    assert np.array_equal(s[:, 2], s[:, 0][..., ::-1, ::-1])                # symmetry 2 = rotation by 180
    assert np.array_equal(s[:, 7], np.swapaxes(s[:, 0], -1, -2))            # symmetry 7 = transpose
    d = dists.reshape(-1, 8, 65)
    assert np.array_equal(d[:, 4, :64].reshape(-1, 8, 8), d[:, 0, :64].reshape(-1, 8, 8)[..., ::-1])
    assert np.array_equal(d[:, :, 64], np.repeat(d[:, :1, 64], 8, 1))
    o = outcomes.reshape(-1, 8)
    assert (o == o[:, :1]).all()
    colour = s[:, 0, 2, 0, 0]                                                # 1.0 when player ZERO moved
    assert set(np.unique(colour)) <= {0.0, 1.0}
