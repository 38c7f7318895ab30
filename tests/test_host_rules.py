"""CPU tests of the product's bitboard rules and RNG/math contract: the HOST
instantiation of sprl_b200/csrc/{games,rng}.cuh (tests/hostcheck/hostcheck.cu,
compiled here with nvcc as host code) against the oracle.  The device
instantiation of the same code is covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so")
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets", "-diag-suppress", "20011", "-diag-suppress", "20014",
                    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wno-unknown-pragmas", "-shared", "-o", out,
                    os.path.join(HERE, "hostcheck", "hostcheck.cu")], check=True, capture_output=True)
    lib = C.CDLL(out)
    lib.hostcheck_rollout.restype = C.c_int64
    lib.hostcheck_perft.restype = C.c_int64
    lib.hostcheck_philox.restype = C.c_uint32
    lib.hostcheck_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    lib.hostcheck_powf.restype = C.c_float
    lib.hostcheck_powf.argtypes = [C.c_float, C.c_float]
    lib.hostcheck_expf.restype = C.c_float
    lib.hostcheck_expf.argtypes = [C.c_float]
    lib.hostcheck_gamma.restype = C.c_float
    return lib


def _rollout(lib, game, seed, first, n):
    gi = O.game_info(game)
    cap = n * 400
    r = dict(game_steps=np.zeros(n, np.int32), cells=np.zeros((cap, gi.cells), np.int8), player=np.zeros(cap, np.int8),
             terminal=np.zeros(cap, np.int8), winner=np.zeros(cap, np.int8), mask=np.zeros((cap, gi.actions), np.int8),
             action=np.zeros(cap, np.int32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cnt = lib.hostcheck_rollout(C.c_int(game), C.c_uint64(seed), C.c_uint64(first), C.c_int(n), C.c_int64(cap),
                                p(r["game_steps"]), p(r["cells"]), p(r["player"]), p(r["terminal"]), p(r["winner"]),
                                p(r["mask"]), p(r["action"]))
    assert cnt >= 0
    return {k: (v if k == "game_steps" else v[:cnt]) for k, v in r.items()}


@pytest.mark.parametrize("game,n", [(O.OG_OTHELLO, 300), (O.OG_C4, 400), (O.OG_GO7, 600), (O.OG_GO9, 150)])
def test_bitboard_rollouts_match_oracle(hc, game, n):
    a = _rollout(hc, game, 5, 0, n)
    b = O.rollout(game, 5, 0, n)
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("game,depth", [(O.OG_OTHELLO, 7), (O.OG_C4, 7), (O.OG_GO7, 4), (O.OG_GO9, 3)])
def test_bitboard_perft_matches_oracle(hc, game, depth):
    out = C.c_uint64()
    assert hc.hostcheck_perft(C.c_int(game), C.c_int(depth), C.byref(out)) == 0
    assert out.value == O.perft(game, depth)


def test_contract_restatement_bit_identical(hc):
    """The product's restatement of the RNG/math contract equals the oracle's, bit for bit."""
    ol = O.lib()
    for seed, game, ctr in [(0, 0, 0), (1, 2, 3), (2**63 + 5, 2**40 + 1, 2**33 + 9), (7, 2**64 - 1, 123456789)]:
        assert hc.hostcheck_philox(seed, game, ctr) == ol.oracle_philox(seed, game, ctr)
    xs = np.concatenate([np.linspace(0, 1, 3001), np.random.RandomState(0).rand(3000)]).astype(np.float32)
    for e in (0.98, 10.0):
        a = np.array([hc.hostcheck_powf(float(x), e) for x in xs], np.float32)
        b = np.array([ol.oracle_det_powf(float(x), e) for x in xs], np.float32)
        assert np.array_equal(a.view(np.int32), b.view(np.int32))
    ls = np.random.RandomState(1).uniform(-40, 40, 4000).astype(np.float32)
    a = np.array([hc.hostcheck_expf(float(x)) for x in ls], np.float32)
    b = np.array([ol.oracle_det_expf(float(x)) for x in ls], np.float32)
    assert np.array_equal(a.view(np.int32), b.view(np.int32))
    # gamma draws through Dirichlet with n=1 slot is always 1; compare raw gammas via two-slot ratios
    for alpha in (0.2, 0.3, 0.5, 1.0, 2.5):
        for g in range(50):
            ctr = C.c_uint64(0)
            g0 = hc.hostcheck_gamma(C.c_uint64(3), C.c_uint64(g), C.byref(ctr), C.c_float(alpha))
            g1 = hc.hostcheck_gamma(C.c_uint64(3), C.c_uint64(g), C.byref(ctr), C.c_float(alpha))
            out = np.zeros(2, np.float32)
            octr = C.c_uint64()
            ol.oracle_dirichlet(C.c_uint64(3), C.c_uint64(g), C.c_uint64(0), C.c_float(alpha),
                                out.ctypes.data_as(C.c_void_p), C.c_int(2), C.byref(octr))
            s = np.float32(np.float32(0) + np.float32(g0)) + np.float32(g1)
            norm = np.float32(1) / s
            want = np.array([np.float32(g0) * norm, np.float32(g1) * norm], np.float32)
            assert octr.value == ctr.value
            assert np.array_equal(out.view(np.int32), want.view(np.int32))
