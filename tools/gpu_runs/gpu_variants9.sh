run() { python bench.py --no-e2e --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sims/s %.2f M  k_round %.4f ms  evalnet %.4f ms' % (d['value']/1e6, d['roofline_search']['launch_ms'], d['roofline']['launch_ms']))"; }
echo "=== default"; tools/build_search_variant.sh && run
echo "=== cold calls"; tools/build_search_variant.sh -DSPRL_SEARCH_COLD_CALLS 2>&1 | grep -i "error" ; run
echo "=== default again"; tools/build_search_variant.sh && run
