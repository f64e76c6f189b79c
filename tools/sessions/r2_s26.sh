#!/bin/bash
# round 2, GPU session 26: grid filter with the powered operator (MFS_BF_FLAG_POWER_OPERATOR): parity, per-record probe, bench secondary leg
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_brute_force.py tests/test_host_api.py -q -m gpu --maxfail=10 --timeout 600 > $O/r2_s26_pytest.log 2>&1
echo "pytest exit $?"; tail -5 $O/r2_s26_pytest.log
timeout 600 python tools/bf_per_record_probe.py 8 100 > $O/r2_s26_bf_per_record.log 2>&1; cat $O/r2_s26_bf_per_record.log
timeout 1500 python bench.py --no-e2e --no-cpu > $O/r2_s26_bench.json 2> $O/r2_s26_bench.err; echo "bench exit $?"; tail -3 $O/r2_s26_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_s26_bench.json') if l.startswith('{')][-1])
g=d['secondary']['grid_filter']
print('literal', g['value'], g['ms'], g['roofline']['frac']); print('power', g['power_operator'])
PY
