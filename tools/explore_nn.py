"""Exploration (not part of the product): how fast can the reference's traced fp32 network run
on a B200 through LibTorch / cuDNN, without leaving fp32?  Variants: cuDNN autotune
(cudnn.benchmark), channels_last, frozen module (conv+BN folded)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200.network import make_network, trace_network

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ref_net = make_network("othello", 0)
x_cpu = (torch.rand(256, 3, 8, 8) > 0.5).float()
with torch.no_grad():
    ref_logits, ref_value = ref_net(x_cpu)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for frozen in (False, True):
        for cl in (False, True):
            net = trace_network(make_network("othello", 0), dev)
            if frozen:
                net = torch.jit.freeze(net)
            x = (torch.rand(B, 3, 8, 8, device=dev) > 0.5).float()
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
                if not frozen:
                    net = net.to(memory_format=torch.channels_last)
            try:
                with torch.no_grad():
                    ms = timeit(lambda: net(x))
                    lg, v = net(x_cpu.to(dev).contiguous(memory_format=torch.channels_last) if cl else x_cpu.to(dev))
                err = (lg.cpu() - ref_logits).abs().max().item()
                print(f"B={B} cudnn.benchmark={bench} frozen={frozen} channels_last={cl}: {ms:.3f} ms  {B / ms * 1e3 / 1e6:.2f} M evals/s  "
                      f"{B * 19.1e6 / ms / 1e9:.1f} TFLOP/s  max|dlogit| vs CPU fp32 = {err:.2e}", flush=True)
            except Exception as ex:
                print(f"B={B} bench={bench} frozen={frozen} cl={cl}: failed {type(ex).__name__}: {str(ex)[:200]}", flush=True)

# TF32 for context only (not used: the reference evaluates in fp32)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
net = trace_network(make_network("othello", 0), dev)
x = (torch.rand(B, 3, 8, 8, device=dev) > 0.5).float()
with torch.no_grad():
    ms = timeit(lambda: net(x))
    lg, v = net(x_cpu.to(dev))
print(f"[context] TF32 allowed: {ms:.3f} ms {B / ms * 1e3 / 1e6:.2f} M evals/s  max|dlogit| = {(lg.cpu() - ref_logits).abs().max().item():.2e}", flush=True)
