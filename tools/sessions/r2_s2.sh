#!/bin/bash
# round 2, GPU session 2: full GPU test suite, ncu captures (1-D N=8, 2-D N=5), ys layout A/B, small-scale bench, memcheck
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=8 > $O/r2_s2_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s2_pytest.log
tail -15 $O/r2_s2_pytest.log
timeout 300 python tools/ys_layout_ab.py > $O/r2_s2_ys_layout.log 2>&1; cat $O/r2_s2_ys_layout.log
timeout 900 python bench.py --steps 2 --warmup 3 --secondary-scale 0.1 > $O/r2_s2_bench.json 2> $O/r2_s2_bench.err; echo "bench exit $?"; tail -3 $O/r2_s2_bench.err; head -c 9000 $O/r2_s2_bench.json
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s2_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v6 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s2_ncu1.log 2>&1
tail -2 $O/r2_s2_profile_case.log
timeout 600 python tools/nd_profile_case.py > $O/r2_s2_nd_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N5_v6 -f python tools/nd_profile_case.py > $O/r2_s2_ncu2.log 2>&1
tail -2 $O/r2_s2_nd_case.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_cases.py > $O/r2_s2_memcheck.log 2>&1; echo "memcheck exit $?" >> $O/r2_s2_memcheck.log; tail -6 $O/r2_s2_memcheck.log
