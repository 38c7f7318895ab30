set -x
timeout 200 python tools/diag_rows.py
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from sprl_b200 import capi, selfplay as SP
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network
outs = []
for rep in range(2):
    ev = EvalNet(make_network("othello", 0), device=0)
    with SP.Engine(capi.GAME_OTHELLO, capi.EVAL_EXTERNAL, seed=5, sims=100, max_batch=8, max_queue=4, num_slots=300, max_games=600) as eng:
        eng.attach_evalnet(ev, use_cuda_graph=True)
        outs.append(eng.run_iteration(600))
print("two runs identical:", all(np.array_equal(a, b) for a, b in zip(*outs)), outs[0][0].shape)
PY
