timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|dlogit" | tail -2
timeout 120 python tools/check_evalnet.py 65536 2 2>&1 | grep "forward B" | tail -1
SPRL_EVALNET_DEBUG=6 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B" | tail -1
