set -x
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|dlogit" | tail -2
timeout 600 python - <<'PY'
import json, time, torch, sys
sys.path.insert(0, '.')
from sprl_b200 import capi, selfplay as SP
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network
net = make_network("go9", 0)
ev = EvalNet(net, device=0, rows=9, cols=9)
x = (torch.rand(8192, 17, 9, 9) > 0.5).float().cuda()
for _ in range(3): ev(x)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(10): ev(x)
torch.cuda.synchronize(); print("go9 evalnet 8192 leaves: %.3f ms" % ((time.time() - t0) * 100))
with SP.Engine(capi.GAME_GO9, capi.EVAL_EXTERNAL, seed=0, sims=400, max_batch=16, max_queue=8, dir_eps=0.25, dir_alpha=0.2,
               num_slots=1024, max_games=8192) as eng:
    eng.attach_evalnet(ev, use_cuda_graph=True)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.begin_iteration(0, 8192)
    eng._capture()
    g = eng._nn["graph"]
    for _ in range(100): g.replay()
    torch.cuda.synchronize(); eng.reset_stats(); t0 = time.time()
    for _ in range(300): g.replay()
    torch.cuda.synchronize(); dt = time.time() - t0
    st = eng.stats(); print("go9 selfplay: %.2f M sims/s, %.1f moves/s, round %.2f ms, failed %s" % (st["sims"] / dt / 1e6, st["moves"] / dt, dt / 300 * 1e3, eng.poll()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.round()
    e1.record(); torch.cuda.synchronize(); print("go9 k_round alone: %.3f ms" % (e0.elapsed_time(e1) / 20))
ev.status()
PY
