#!/bin/bash
# round 2, GPU session 38: full suite + smoke + bench + N sweep with the final library (6 CTAs per SM up to N = 8, 5 at N = 9, 10, 4 up to N = 14; step barrier from N = 9)
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 > $O/r2_s38_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s38_pytest.log
tail -4 $O/r2_s38_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_s38_smoke.log 2>&1; tail -2 $O/r2_s38_smoke.log
timeout 1500 python bench.py > $O/r2_bench_line_v8.json 2> $O/r2_s38_bench.err; echo "bench exit $?"; tail -3 $O/r2_s38_bench.err; head -c 300 $O/r2_bench_line_v7.json; echo
timeout 600 python tools/time_profile.py > $O/r2_time_profile_N_sweep.md 2> $O/r2_s38_tp.err; tail -16 $O/r2_time_profile_N_sweep.md
