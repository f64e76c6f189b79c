#!/bin/bash
# round 2, GPU session 20: full suite + smoke + bench + N sweep with the step-barrier policy (1-D value kernel, gradient kernel) and the final 2-D kernel
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 > $O/r2_s20_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s20_pytest.log
tail -4 $O/r2_s20_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_s20_smoke.log 2>&1; tail -2 $O/r2_s20_smoke.log
timeout 1500 python bench.py > $O/r2_bench_line_v5.json 2> $O/r2_s20_bench.err; echo "bench exit $?"; tail -3 $O/r2_s20_bench.err; head -c 400 $O/r2_bench_line_v5.json; echo
timeout 600 python tools/time_profile.py > $O/r2_time_profile_N_sweep.md 2> $O/r2_s20_tp.err; tail -16 $O/r2_time_profile_N_sweep.md
timeout 300 python tools/grad_probe.py > $O/r2_s20_grad.log 2>&1; cat $O/r2_s20_grad.log
