"""A/B of the two conv-tower kernels on the GPU: agreement of the outputs and CUDA-event timing.
    python tools/check_resident.py [rows ...]        (default 65536 61960 4096)"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200 import capi
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network

import threading
import time
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler:
    """SM clock and board power while a timing loop runs (NVML, every 5 ms)."""
    def __enter__(self):
        self.clk, self.pw, self.stop = [], [], False
        def run():
            while not self.stop:
                self.clk.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
                self.pw.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3)
                time.sleep(0.005)
        self.t = threading.Thread(target=run)
        self.t.start()
        return self
    def __exit__(self, *a):
        self.stop = True
        self.t.join()
    def summary(self):
        c, p = sorted(self.clk), sorted(self.pw)
        return f"sm clock median {c[len(c) // 2]} MHz (min {c[0]}), power median {p[len(p) // 2]:.0f} W (max {p[-1]:.0f}) over {len(c)} samples"


REPS = int(os.environ.get("REPS", "20"))
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
        reps = REPS
        with Sampler() as smp:
            e0.record()
            for _ in range(reps):
                ev(x)
            e1.record()
            torch.cuda.synchronize()
        if reps >= 100:
            print(f"  {name}: {smp.summary()}", flush=True)
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
