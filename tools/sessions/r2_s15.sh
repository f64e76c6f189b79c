#!/bin/bash
# round 2, GPU session 15: full suite + smoke + bench with the final 2-D kernel (phase barriers, 8 x 2 CTA shape, MEANVAR for the 2-D filter)
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 > $O/r2_s15_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s15_pytest.log
tail -6 $O/r2_s15_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_s15_smoke.log 2>&1; tail -2 $O/r2_s15_smoke.log
timeout 1500 python bench.py > $O/r2_bench_line_v4.json 2> $O/r2_s15_bench.err; echo "bench exit $?"; tail -3 $O/r2_s15_bench.err; head -c 600 $O/r2_bench_line_v4.json
