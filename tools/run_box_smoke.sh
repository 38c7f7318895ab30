#!/bin/bash
# two GPUs, one OTHWorker each (iteration 0, uniform evaluator): both task directories get their three .npy files
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=$(mktemp -d); cd "$WORK"
export SPRL_RUN_NAME=box SPRL_NUM_GROUPS=2 SPRL_NUM_ITERS=1 SPRL_INIT_NUM_GAMES=32 SPRL_INIT_UCT_TRAVERSALS=48 SPRL_INIT_MAX_BATCH_SIZE=8 SPRL_INIT_MAX_QUEUE_SIZE=4
"$ROOT/sprl_b200/host/run_box.sh" OTHWorker 4 8 2
python - <<PY
import numpy as np
for task, group in ((4, 1), (5, 1)):
    b = f"data/games/box/{group}/{task}/box_iteration_0"
    s, d, o = np.load(b + "_states.npy"), np.load(b + "_distributions.npy"), np.load(b + "_outcomes.npy")
    assert s.shape[1:] == (3, 8, 8) and d.shape == (s.shape[0], 65) and o.shape == (s.shape[0],) and s.shape[0] >= 32 * 8 * 20
    print("task", task, "ok:", s.shape[0], "samples")
a = np.load("data/games/box/1/4/box_iteration_0_states.npy"); b = np.load("data/games/box/1/5/box_iteration_0_states.npy")
assert a.shape != b.shape or not np.array_equal(a, b), "the two tasks must play different game streams"
PY
echo "run_box smoke ok"
