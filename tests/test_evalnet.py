"""The library's evaluator network (csrc/evalnet.cu, sprl_evalnet_*) against the fp64 / fp32
PyTorch forward of the same `BasicGridNetwork` (the module the reference traces,
src/networks/grid_networks.py:30-80).  Floating-point kernel: tolerance, stated per test."""
import ctypes as C

import numpy as np
import pytest
import torch

from sprl_b200 import capi
from sprl_b200.evalnet import EvalNet, network_params
from sprl_b200.network import BasicGridNetwork, make_network


def randomized(net, seed):
    """Non-trivial BatchNorm statistics so that the folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.rand(mod.running_mean.shape, generator=g) - 0.5)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 1.5 + 0.5)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.rand(mod.bias.shape, generator=g) - 0.5)
    return net


# ------------------------------------------------------------------ CPU: argument validation
def test_evalnet_rejects_unsupported_shapes():
    lib = capi.load()
    h = C.c_void_p()
    # more than 128 cells do not fit one MMA tile
    p, keep = network_params(make_network("go9", 0).state_dict(), rows=12, cols=12)
    assert lib.sprl_evalnet_create(0, C.byref(p), C.byref(h)) == capi.SPRL_E_INVALID and not h
    assert b"8x8" in lib.sprl_last_error()
    # wrong tower width
    p, keep = network_params(BasicGridNetwork(8, 8, 65, 1, 1, 32).state_dict())
    assert lib.sprl_evalnet_create(0, C.byref(p), C.byref(h)) == capi.SPRL_E_INVALID
    # null pointer inside the parameter block
    p, keep = network_params(make_network("othello", 0).state_dict())
    p.policy_fc_w = None
    assert lib.sprl_evalnet_create(0, C.byref(p), C.byref(h)) == capi.SPRL_E_INVALID
    assert lib.sprl_evalnet_create(0, None, C.byref(h)) == capi.SPRL_E_INVALID
    assert lib.sprl_evalnet_forward(None, None, 4, None, None, None) == capi.SPRL_E_INVALID


def test_network_params_follow_the_state_dict():
    net = make_network("othello", 3)
    p, keep = network_params(net.state_dict())
    assert (p.rows, p.cols, p.in_planes, p.channels, p.blocks, p.actions) == (8, 8, 3, 64, 2, 65)
    assert (p.policy_channels, p.value_channels, p.value_hidden) == (2, 1, 64)
    w = np.ctypeslib.as_array(C.cast(p.tower[3].weight, C.POINTER(C.c_float)), shape=(64 * 64 * 9,))
    assert np.array_equal(w, net.residual_blocks[1].conv2.weight.detach().numpy().reshape(-1))


# ------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("batch,blocks,planes", [(1, 2, 3), (2, 2, 3), (257, 2, 3), (4099, 2, 3), (64, 0, 3), (130, 6, 3), (96, 1, 17)])
def test_evalnet_matches_fp64_forward(batch, blocks, planes):
    """|dlogit| and |dvalue| <= 2e-6 against the fp64 forward (the fp32 PyTorch forward itself is
    within ~2e-7); odd batches, tower depths 0..6 and multi-k-step stems (17 planes)."""
    torch.manual_seed(blocks * 7 + planes)
    net = randomized(BasicGridNetwork(8, 8, 65, (planes - 1) // 2, blocks, 64).eval(), 5)
    x = (torch.rand(batch, planes, 8, 8) > 0.5).float()
    with torch.no_grad():
        want_l, want_v = net.double()(x.double())
        net.float()
    ev = EvalNet(net, device=0)
    got_l, got_v = ev(x.cuda())
    ev.status()
    assert got_l.shape == (batch, 65) and got_v.shape == (batch, 1)
    assert (got_l.cpu().double() - want_l).abs().max().item() <= 2e-6
    assert (got_v.cpu().double() - want_v).abs().max().item() <= 2e-6
    ev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,rows,cols,batch", [("c4", 6, 7, 131), ("go7", 7, 7, 77), ("go9", 9, 9, 203)])
def test_evalnet_smaller_boards(kind, rows, cols, batch):
    """Connect Four (6x7, 7 actions) and Go 7x7 (17 planes, 6 blocks, 50 actions) on the 8x8 lattice; Go 9x9 (82 actions)
    on the linear lattice, one board per tile, neighbours across warp boundaries exchanged through shared memory."""
    net = randomized(make_network(kind, 3), 9)
    planes = net.conv.in_channels
    x = (torch.rand(batch, planes, rows, cols) > 0.5).float()
    with torch.no_grad():
        want_l, want_v = net.double()(x.double())
        net.float()
    ev = EvalNet(net, device=0, rows=rows, cols=cols)
    got_l, got_v = ev(x.cuda())
    ev.status()
    assert (got_l.cpu().double() - want_l).abs().max().item() <= 2e-6
    assert (got_v.cpu().double() - want_v).abs().max().item() <= 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("kind,rows,cols,batch,phases", [("othello", 8, 8, 4099, 2), ("othello", 8, 8, 5, 2), ("c4", 6, 7, 1031, 2),
                                                         ("go7", 7, 7, 300, 7), ("go9", 9, 9, 203, 7)])
def test_resident_and_streaming_kernels_agree(kind, rows, cols, batch, phases):
    """The resident-weight kernel (CTA pairs, one launch per residual block, activations through HBM) accumulates the
    conv tower in the same order as the streaming kernel; its 1x1 head convolutions are fp32 FMAs inside the last conv
    epilogue instead of split-fp16 MMAs, so the two agree to rounding (1e-6 on logits and value, both within 2e-6 of the
    fp64 forward in the tests above) -- two boards per tile on the 8x8 lattice, one on the linear lattice (Go 9x9).  The
    resident path itself is bit-identical from one grid size to another."""
    net = randomized(make_network(kind, 2), 6)
    x = (torch.rand(batch, net.conv.in_channels, rows, cols) > 0.5).float().cuda()
    ev = EvalNet(net, device=0, rows=rows, cols=cols)
    assert ev.phases == phases                               # AUTO picks the resident kernel for these boards
    l_res, v_res = ev(x)
    ev.set_path(capi.EVALNET_PATH_STREAMING)
    assert ev.phases == 0
    l_str, v_str = ev(x)
    ev.set_path(capi.EVALNET_PATH_RESIDENT)
    l_again, v_again = ev(x[: batch // 2 + 1])              # another grid size on the resident path
    ev.status()
    assert (l_res - l_str).abs().max().item() <= 1e-6 and (v_res - v_str).abs().max().item() <= 1e-6
    assert torch.equal(l_again, l_res[: batch // 2 + 1]) and torch.equal(v_again, v_res[: batch // 2 + 1])
    ev.close()


@pytest.mark.gpu
def test_evalnet_real_valued_inputs():
    """The reference's planes are 0/1, but the kernel splits any fp32 input into hi + lo."""
    net = randomized(make_network("othello", 1), 2)
    x = torch.rand(64, 3, 8, 8)
    with torch.no_grad():
        want = net(x)[0]
    got = EvalNet(net, device=0)(x.cuda())[0].cpu()
    assert (got - want).abs().max().item() <= 2e-6


@pytest.mark.gpu
def test_evalnet_update_in_place_and_batches_in_a_row():
    a, b = randomized(make_network("othello", 1), 1), randomized(make_network("othello", 2), 2)
    x = (torch.rand(300, 3, 8, 8) > 0.5).float()
    ev = EvalNet(a, device=0)
    xg = x.cuda()
    la = ev(xg)[0].cpu()
    ev.update(b)
    lb = ev(xg)[0].cpu()
    with torch.no_grad():
        assert (la - a(x)[0]).abs().max().item() <= 2e-6 and (lb - b(x)[0]).abs().max().item() <= 2e-6
    assert (la - lb).abs().max().item() > 1e-3
    for n in (1, 7, 300, 33):                      # different grid sizes back to back
        assert torch.equal(ev(xg[:n])[0].cpu(), lb[:n])
    assert ev.status() >= 6
    with pytest.raises(capi.SprlError):
        ev.update(make_network("go7", 0))          # shape change is refused


@pytest.mark.gpu
def test_selfplay_with_evalnet_tracks_libtorch():
    """Same games through both evaluators: the tree statistics are discontinuous in the priors, so
    compare what is continuous -- the first move's root priors (one forward) -- and that whole
    games finish with well-formed samples."""
    from sprl_b200 import selfplay as SP
    from sprl_b200.network import trace_network
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = randomized(make_network("othello", 4), 4)
    kw = dict(seed=8, sims=48, max_batch=8, max_queue=4, num_slots=16, max_games=16, record_stats=1, add_noise=0)
    out = {}
    for kind in ("evalnet", "libtorch"):
        with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, **kw) as eng:
            if kind == "evalnet":
                ev = EvalNet(net, device=0)
                eng.attach_evalnet(ev, use_cuda_graph=True)
            else:
                eng.attach_network(trace_network(net, torch.device("cuda", 0)), use_cuda_graph=False)
            states, dists, outcomes = eng.run_iteration(16)
            ms = eng.move_stats(16)
            out[kind] = (ms, states, dists, outcomes)
        if kind == "evalnet":
            ev.status()
    (ma, sa, da, oa), (mb, sb, db, ob) = out["evalnet"], out["libtorch"]
    first_a = np.concatenate([[0], np.cumsum(ma["game_moves"])[:-1]])
    first_b = np.concatenate([[0], np.cumsum(mb["game_moves"])[:-1]])
    assert np.allclose(ma["move_P"][first_a], mb["move_P"][first_b], atol=1e-5)
    for s, d, o in ((sa, da, oa), (sb, db, ob)):
        assert s.shape[0] == d.shape[0] == o.shape[0] and s.shape[0] % 8 == 0
        assert np.allclose(d.sum(1), 1.0, atol=1e-5) and set(np.unique(o)) <= {-1.0, 0.0, 1.0}


@pytest.mark.gpu
def test_evalnet_outputs_do_not_depend_on_the_row():
    """The engine hands leaf rows out in no fixed order (one atomicAdd per tree), so reproducible self-play needs
    a board's outputs to be bit-identical wherever it sits in the batch, whoever its tile partner is."""
    net = randomized(make_network("othello", 1), 4)
    ev = EvalNet(net, device=0)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(1501, 3, 8, 8, generator=g) > 0.5).float().cuda()
    perm = torch.randperm(x.shape[0], generator=g).cuda()
    l0, v0 = ev(x)
    l1, v1 = ev(x[perm])
    assert torch.equal(l0[perm], l1) and torch.equal(v0[perm], v1)
    l2, v2 = ev(x[:77])                                       # a different grid size, other CTAs
    assert torch.equal(l0[:77], l2) and torch.equal(v0[:77], v2)
    ev.status()


@pytest.mark.gpu
def test_selfplay_through_the_evaluator_is_reproducible():
    """Same seed, same network: identical samples, although leaf rows are handed out by atomics in a different order."""
    import numpy as np
    from sprl_b200 import selfplay as SP
    outs = []
    for _ in range(2):
        ev = EvalNet(make_network("othello", 0), device=0)
        with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=5, sims=64, max_batch=8, max_queue=4, num_slots=96,
                       max_games=160) as eng:
            eng.attach_evalnet(ev, use_cuda_graph=True)
            outs.append(eng.run_iteration(160))
        ev.close()
    assert outs[0][0].shape[0] > 160 * 8 * 20
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_evalnet_reports_out_of_range_activations():
    """The fp16 split covers inputs and activations up to 4,094; beyond that the forward is reported, not silently clamped."""
    net = make_network("othello", 2)
    ev = EvalNet(net, device=0)
    x = torch.full((4, 3, 8, 8), 5000.0).cuda()
    ev(x)
    with pytest.raises(capi.SprlError) as err:
        ev.status()
    assert "range" in str(err.value)
    ev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gain,ok", [(6.0, True), (40.0, False)])
def test_evalnet_large_batchnorm_scales(gain, ok):
    """A trained network can carry BatchNorm scales well above the random-init ones.  With every BN weight multiplied by
    `gain` the activations of the second block reach the hundreds (gain 6: 513, still inside the fp16 split's range of
    4,094) or leave the range (gain 40: the forward must be REPORTED by sprl_evalnet_status, on the resident and on the
    streaming kernel alike).  Accuracy in the first case: the hi/lo split carries 22 significant bits per operand, so the
    error is bounded RELATIVE TO THE LARGEST INTERMEDIATE ACTIVATION -- 1e-6 of it here (measured 4.5e-7: 2.3e-4 on logits
    of up to 88 built from activations of up to 513; PyTorch's fp32 CPU forward of the same network is 3.4e-5 from fp64)."""
    net = randomized(make_network("othello", 2), 9)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.mul_(gain)
    x = (torch.rand(257, 3, 8, 8) > 0.5).float()
    acts = []
    hooks = [mod.register_forward_hook(lambda _m, _i, out: acts.append(out.abs().max().item()))
             for mod in net.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        want_l, want_v = net.double()(x.double())
        net.float()
    for h in hooks:
        h.remove()
    for path in (capi.EVALNET_PATH_RESIDENT, capi.EVALNET_PATH_STREAMING):
        ev = EvalNet(net, device=0)                              # (the out-of-range report is sticky: one evaluator per path)
        ev.set_path(path)
        got_l, got_v = ev(x.cuda())
        if ok:
            ev.status()
            scale = max(acts)
            assert scale > 300.0                                 # the test is about large activations
            assert (got_l.cpu().double() - want_l).abs().max().item() <= 1e-6 * scale
            assert (got_v.cpu().double() - want_v).abs().max().item() <= 1e-5
        else:
            with pytest.raises(capi.SprlError) as err:
                ev.status()
            assert "range" in str(err.value)
        ev.close()


@pytest.mark.gpu
def test_evalnet_single_pass_fp16_is_opt_in_and_labelled():
    """SPRL_EVALNET_PRECISION_FP16: one fp16 MMA per product instead of the three of the hi/lo split.  Opt-in, resident
    kernel only; accuracy of fp16 operands with fp32 accumulation (measured here, bound 5e-3 on logits of O(1)); switching
    back restores the default path's bits; the streaming kernel refuses the mode."""
    net = randomized(make_network("othello", 2), 4)
    x = (torch.rand(515, 3, 8, 8) > 0.5).float()
    with torch.no_grad():
        want_l, want_v = net.double()(x.double())
        net.float()
    ev = EvalNet(net, device=0)
    l32, v32 = ev(x.cuda())
    ev.set_precision(capi.EVALNET_PRECISION_FP16)
    l16, v16 = ev(x.cuda())
    l16b, _ = ev(x[:77].cuda())                                   # deterministic and independent of the batch
    ev.status()
    e32 = (l32.cpu().double() - want_l).abs().max().item()
    e16 = (l16.cpu().double() - want_l).abs().max().item()
    assert e32 <= 2e-6
    assert 1e-5 < e16 <= 5e-3, e16                                 # a different, coarser arithmetic -- and only when asked for
    assert (v16.cpu().double() - want_v).abs().max().item() <= 5e-3
    assert torch.equal(l16b, l16[:77])
    ev.set_precision(capi.EVALNET_PRECISION_FP32_SPLIT)
    l_again, v_again = ev(x.cuda())
    assert torch.equal(l_again, l32) and torch.equal(v_again, v32)
    ev.set_precision(capi.EVALNET_PRECISION_FP16)
    ev.set_path(capi.EVALNET_PATH_STREAMING)
    with pytest.raises(capi.SprlError):
        ev(x.cuda())
    ev.close()
