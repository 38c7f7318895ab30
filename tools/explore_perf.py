"""Exploration (not part of the product): time the traced network and the search launch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sprl_b200 import capi, selfplay as SP
from sprl_b200.network import make_network, trace_network

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
net = trace_network(make_network("othello", 0), dev)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    for B in (1024, 4096, 16384, 65536):
        x = torch.zeros(B, 3, 8, 8, device=dev)
        with torch.no_grad():
            ms = timeit(lambda: net(x))
        print(f"net fp32 tf32={tf32} B={B}: {ms:.3f} ms  {B/ms*1e3/1e6:.2f} M evals/s  {B*19.1e6/ms/1e9:.1f} TFLOP/s", flush=True)
torch.backends.cudnn.allow_tf32 = False
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for B in (16384, 65536):
        x = torch.zeros(B, 3, 8, 8, device=dev)
        ms = timeit(lambda: net(x))
        print(f"net bf16 autocast B={B}: {ms:.3f} ms {B/ms*1e3/1e6:.2f} M evals/s", flush=True)

# search-only throughput with the device evaluators
for ev, name in ((capi.EVAL_UNIFORM, "uniform"), (capi.EVAL_HASHNET, "hashnet")):
    for G in (1024, 4096, 16384):
        with SP.Engine(capi.GAME_OTHELLO, ev, sims=400, max_batch=8, max_queue=4, num_slots=G, max_games=G, rounds_per_launch=16) as eng:
            t = time.time(); eng.run_iteration(G, collect=False); dt = time.time() - t
            st = eng.stats()
            print(f"search {name} G={G}: {dt:.2f}s {st['sims']/dt/1e6:.2f} M sims/s {st['moves']/dt:.0f} moves/s depth {st['depth_sum']/st['sims']:.2f} L {st['legal_sum']/max(1,st['nodes_visited']):.2f} hw {st['units_high_water']}/{st['units_per_tree']} launches {st['launches']}", flush=True)

# external evaluator, with and without CUDA graph
for G, graph in ((1024, False), (1024, True), (4096, True)):
    with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, sims=400, max_batch=8, max_queue=4, num_slots=G, max_games=G) as eng:
        eng.attach_network(net, use_cuda_graph=graph)
        t = time.time(); eng.run_iteration(G, collect=False); torch.cuda.synchronize(); dt = time.time() - t
        st = eng.stats()
        print(f"selfplay net G={G} graph={graph}: {dt:.2f}s {st['sims']/dt/1e6:.3f} M sims/s {st['moves']/dt:.0f} moves/s evals/sim {st['evals']/st['sims']:.3f} launches {st['launches']}", flush=True)
