#!/bin/bash
# Runs the C++ OTHWorker drop-in (sprl_b200/host/bin/OTHWorker) for two iterations on a GPU box:
# iteration 0 with the uniform evaluator, iteration 1 with a traced network written the way the
# reference's controller does; then checks the .npy files with numpy.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=$(mktemp -d)
cd "$WORK"
export SPRL_RUN_NAME=smoke SPRL_NUM_GROUPS=2 SPRL_NUM_ITERS=2
export SPRL_INIT_NUM_GAMES=64 SPRL_INIT_UCT_TRAVERSALS=64 SPRL_INIT_MAX_BATCH_SIZE=1 SPRL_INIT_MAX_QUEUE_SIZE=1
export SPRL_NUM_GAMES=64 SPRL_UCT_TRAVERSALS=100 SPRL_MAX_BATCH_SIZE=8 SPRL_MAX_QUEUE_SIZE=4
mkdir -p data/models/smoke
PYTHONPATH=$ROOT python - <<PY
import torch
from sprl_b200.network import make_network, trace_network
trace_network(make_network("othello", 0), "cpu").save("data/models/smoke/traced_smoke_iteration_0.pt")
PY
BIN=${1:-$ROOT/sprl_b200/host/bin/OTHWorker}
case "$BIN" in /*) ;; *) BIN="$ROOT/$BIN" ;; esac
time "$BIN" 3 4
python - <<PY
import numpy as np
for it in (0, 1):
    b = f"data/games/smoke/1/3/smoke_iteration_{it}"
    s, d, o = np.load(b + "_states.npy"), np.load(b + "_distributions.npy"), np.load(b + "_outcomes.npy")
    assert s.shape[1:] == (3, 8, 8) and d.shape == (s.shape[0], 65) and o.shape == (s.shape[0],), (s.shape, d.shape, o.shape)
    assert s.shape[0] % 8 == 0 and s.shape[0] >= 64 * 8 * 10
    assert np.allclose(d.sum(1), 1, atol=1e-5) and set(np.unique(o)) <= {-1.0, 0.0, 1.0}
    print("iteration", it, "ok:", s.shape[0], "samples, mean outcome", o.mean())
PY
echo "worker smoke ok ($BIN)"
