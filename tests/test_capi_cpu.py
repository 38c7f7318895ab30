"""CPU tests of the boundary: libsprl_b200.so builds, loads and exports every
symbol include/sprl_b200.h declares; without a GPU compute calls fail loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sprl_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sprl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sprl_[a-z0-9_]+)\s*\(", text)) - {"sprl_forward_fn"})


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sprl_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_game_info_matches_reference_constants():
    # games/OthelloNode.hpp:8-11, ConnectFourNode.hpp:8-13, GoNode.hpp:16-22
    want = {capi.GAME_OTHELLO: (8, 8, 64, 65, 1, 8), capi.GAME_C4: (6, 7, 42, 7, 1, 2),
            capi.GAME_GO7: (7, 7, 49, 50, 8, 8), capi.GAME_GO9: (9, 9, 81, 82, 8, 8)}
    for g, w in want.items():
        gi = capi.game_info(g)
        assert (gi.rows, gi.cols, gi.cells, gi.actions, gi.history, gi.nsym) == w
    with pytest.raises(capi.SprlError):
        capi.game_info(17)


def test_default_config_is_the_reference_worker():
    c = capi.default_config(capi.GAME_OTHELLO)     # OTHWorker.cpp:23-28, constants.hpp:6
    assert (c.sims, c.max_batch, c.max_queue) == (8192, 8, 4)
    assert abs(c.dir_eps - 0.25) < 1e-7 and abs(c.dir_alpha - 0.3) < 1e-7 and abs(c.u_weight - 1.1) < 1e-7
    assert c.init_q == capi.INITQ_PARENT and c.add_noise == 1 and c.use_sym == 1
    c = capi.default_config(capi.GAME_C4)          # C4Worker.cpp:22-27
    assert (c.sims, c.max_batch, c.max_queue) == (512, 8, 4) and abs(c.dir_alpha - 0.5) < 1e-7
    c = capi.default_config(capi.GAME_GO7)         # GoWorker.cpp:22-27
    assert (c.sims, c.max_batch, c.max_queue) == (32768, 16, 8) and abs(c.dir_alpha - 0.2) < 1e-7


def test_no_gpu_fails_loudly():
    lib = capi.load()
    if lib.sprl_device_count() > 0:
        pytest.skip("a GPU is present")
    count, ms = C.c_uint64(), C.c_float()
    rc = lib.sprl_env_perft(0, capi.GAME_OTHELLO, 3, C.byref(count), C.byref(ms))
    assert rc == capi.SPRL_E_NOGPU
    assert b"no CPU path" in lib.sprl_last_error()
    cfg = capi.default_config(capi.GAME_OTHELLO)
    h = C.c_void_p()
    assert lib.sprl_create(C.byref(cfg), C.byref(h)) == capi.SPRL_E_NOGPU and not h


def test_null_handles_and_bad_options_are_rejected():
    """Argument errors are reported before any device work (so they can be checked without a GPU)."""
    lib = capi.load()
    agents = (capi.AgentConfig * 2)()
    assert lib.sprl_match_begin(None, agents, 0, 4) == capi.SPRL_E_INVALID
    assert lib.sprl_run_match(None, agents, 0, 4, None, None) == capi.SPRL_E_INVALID
    wins, draws = (C.c_int64 * 2)(), C.c_int64()
    assert lib.sprl_match_results(None, 0, None, None, None, wins, C.byref(draws)) == capi.SPRL_E_INVALID
    rows = C.c_void_p()
    assert lib.sprl_eval_rows(None, C.byref(rows)) == capi.SPRL_E_INVALID
    assert lib.sprl_evalnet_forward_counted(None, None, None, 4, None, None, None) == capi.SPRL_E_INVALID
    # evaluator / init-Q ranges (uct/UCTNode.hpp:24-28; networks/OthelloHeuristic.hpp is Othello-only)
    h = C.c_void_p()
    for game, ev, q in [(capi.GAME_C4, capi.EVAL_OTHELLO_HEURISTIC, capi.INITQ_PARENT), (capi.GAME_OTHELLO, 9, capi.INITQ_PARENT),
                        (capi.GAME_OTHELLO, capi.EVAL_UNIFORM, 3)]:
        cfg = capi.default_config(game)
        cfg.evaluator, cfg.init_q = ev, q
        assert lib.sprl_create(C.byref(cfg), C.byref(h)) == capi.SPRL_E_INVALID and not h
    assert capi.INITQ_DROP_PARENT == 2 and capi.EVAL_OTHELLO_HEURISTIC == 3


def test_npy_writer_is_byte_identical_to_reference_layout(tmp_path):
    import oracle_py as O
    from sprl_b200.selfplay import write_npy
    for shape in [(3, 3, 8, 8), (7, 65), (5,), (0, 3, 8, 8), (123456 % 977, 7)]:
        a = np.random.RandomState(0).rand(*shape).astype(np.float32)
        p1, p2 = str(tmp_path / "a.npy"), str(tmp_path / "b.npy")
        write_npy(p1, a)
        O.write_npy(p2, a)
        assert open(p1, "rb").read() == open(p2, "rb").read()
        assert np.array_equal(np.load(p1), a)
    with pytest.raises(capi.SprlError):
        write_npy(str(tmp_path / "nodir" / "x.npy"), np.zeros(3, np.float32))
