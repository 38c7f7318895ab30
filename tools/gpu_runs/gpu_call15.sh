SPRL_EVALNET_TIMING=1 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0\|dlogit" | tail -3
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
python -c "
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network
print(EvalNet(make_network('othello',0)).info(), EvalNet(make_network('go7',0), rows=7, cols=7).info())"
