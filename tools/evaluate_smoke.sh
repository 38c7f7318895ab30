#!/bin/bash
# Runs the C++ Evaluate drop-in (sprl_b200/host/bin/Evaluate, the reference's cpp/src/Evaluate.cpp command line) on a
# GPU box: two traced random-init Connect Four networks, then a network against the uniform evaluator, then the
# Othello heuristic against the uniform evaluator.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=$(mktemp -d)
cd "$WORK"
PYTHONPATH=$ROOT python - <<PY
import torch
from sprl_b200.network import make_network, trace_network
for k in (0, 1):
    trace_network(make_network("c4", k), "cpu").save(f"c4_{k}.pt")
PY
BIN=$ROOT/sprl_b200/host/bin/Evaluate
time "$BIN" c4_0.pt c4_1.pt 64 128 8 4 1 1 1 1 | tee out1.txt
grep -q "Player 0 wins:" out1.txt
time "$BIN" c4_0.pt random 64 128 8 4 1 1 0 0 | tee out2.txt
SPRL_EVALUATOR=libtorch "$BIN" c4_0.pt c4_1.pt 16 64 8 4 1 1 1 1 | tee out3.txt
SPRL_GAME=othello "$BIN" heuristic random 32 100 8 4 1 1 1 1 | tee out4.txt
python - <<PY
import re
for f in ("out1.txt", "out2.txt", "out3.txt", "out4.txt"):
    m = re.search(r"Player 0 wins: (\d+), Player 1 wins: (\d+), Draws: (\d+)", open(f).read())
    n = sum(map(int, m.groups()))
    assert n in (64, 16, 32), (f, n)
    print(f, m.group(0))
PY
echo "evaluate smoke ok"
# the REFERENCE's own Evaluate.cpp, compiled unchanged against the veneer (make -C sprl_b200/host dropin): one game per playGame call
REF_BIN=$ROOT/sprl_b200/host/bin/ref_Evaluate
if [ -x "$REF_BIN" ]; then
  time "$REF_BIN" c4_0.pt c4_1.pt 6 64 8 4 1 1 1 1 | tail -3 | tee out5.txt
  grep -q "Player 0 wins:" out5.txt && echo "reference Evaluate.cpp drop-in ok"
fi
