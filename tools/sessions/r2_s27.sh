#!/bin/bash
# round 2, GPU session 27: full suite + smoke + bench with the final library (all legs) + reference arm
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 > $O/r2_s27_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s27_pytest.log
tail -4 $O/r2_s27_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_s27_smoke.log 2>&1; tail -2 $O/r2_s27_smoke.log
timeout 1500 python bench.py > $O/r2_bench_line_v6.json 2> $O/r2_s27_bench.err; echo "bench exit $?"; tail -3 $O/r2_s27_bench.err; head -c 300 $O/r2_bench_line_v6.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/r2_bench_reference_arm_v6.json 2>> $O/r2_s27_bench.err; tail -c 500 $O/r2_bench_reference_arm_v6.json
