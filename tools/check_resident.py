"""A/B of the two conv-tower kernels on the GPU: bit-equality of the outputs and CUDA-event timing.
    python tools/check_resident.py [rows ...]        (default 65536 61960 4096)"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200 import capi
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network

kind = os.environ.get("KIND", "othello")
shape = {"othello": (8, 8), "c4": (6, 7), "go7": (7, 7), "go9": (9, 9)}[kind]
net = make_network(kind, 0)
planes = net.conv.in_channels
ev = EvalNet(net, device=0, rows=shape[0], cols=shape[1])
print("phases", ev.phases, ev.info(), flush=True)
sizes = [int(a) for a in sys.argv[1:]] or [65536, 61960, 4096]
g = torch.Generator().manual_seed(1)
for n in sizes:
    x = (torch.rand(n, planes, *shape, generator=g) > 0.5).float().cuda()
    out = {}
    for name, path in (("resident", capi.EVALNET_PATH_RESIDENT), ("streaming", capi.EVALNET_PATH_STREAMING)):
        ev.set_path(path)
        l, v = ev(x)
        torch.cuda.synchronize()
        ev.status()
        for _ in range(3):
            ev(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            ev(x)
        e1.record()
        torch.cuda.synchronize()
        out[name] = (l, v, e0.elapsed_time(e1) / reps)
        ev.status()
    (lr, vr, tr), (ls, vs, ts) = out["resident"], out["streaming"]
    dl = (lr - ls).abs().max().item()
    dv = (vr - vs).abs().max().item()
    bad = ((lr != ls).any(1)).nonzero().reshape(-1)
    print(f"rows {n}: resident {tr:.4f} ms  streaming {ts:.4f} ms  (incl. k_heads)  max|dlogit| {dl:.3e} max|dvalue| {dv:.3e} "
          f"rows differing {bad.numel()} first {bad[:8].tolist()}", flush=True)
    if n == sizes[0]:
        with torch.no_grad():
            want = net.double()(x[:512].cpu().double())[0]
            net.float()
        print("  vs fp64 (512 rows): resident", (lr[:512].cpu().double() - want).abs().max().item(),
              "streaming", (ls[:512].cpu().double() - want).abs().max().item(), flush=True)
