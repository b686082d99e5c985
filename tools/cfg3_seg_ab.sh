timeout 300 python -m pytest tests/test_gpu_fullsize.py -k cfg5 -m gpu -q --timeout 300 2>&1 | grep -E "assert [0-9.e-]+ <=|AssertionError|passed|failed" | cut -c1-200
timeout 200 python bench.py --workload cfg5 --steps 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_check'])"
