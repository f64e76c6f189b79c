#!/bin/bash
# round 2, GPU session 9: 2-D kernel with the register-QL fast path (parity + A/B), final 1-D build (rolled atom loops): suite, bench, ncu
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=10 > $O/r2_s9_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s9_pytest.log
tail -6 $O/r2_s9_pytest.log
for a in "5 4736 20" "5 18944 50" "4 18944 50" "3 18944 50" "2 18944 50" "5 18944 50 tme" "5 18944 50 euler"; do
  timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s9_nd_fast.log 2>&1
  MFS_B200_LIB=$PWD/ab/libmfs_ndslow.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s9_nd_r1ql.log 2>&1
done
echo "--- fast path"; cat $O/r2_s9_nd_fast.log; echo "--- round-1 stage D"; cat $O/r2_s9_nd_r1ql.log
timeout 1500 python bench.py > $O/r2_bench_line_v3.json 2> $O/r2_s9_bench.err; echo "bench exit $?"; tail -3 $O/r2_s9_bench.err; head -c 1500 $O/r2_bench_line_v3.json
timeout 600 python tools/nd_profile_case.py 5 4736 20 > $O/r2_s9_nd_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N5_v7 -f python tools/nd_profile_case.py 5 4736 20 > $O/r2_s9_ncu2.log 2>&1
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s9_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v11 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s9_ncu1.log 2>&1
tail -2 $O/r2_s9_profile_case.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_s9_ncu_bench.log 2>&1
timeout 600 python tools/time_profile.py > $O/r2_time_profile_N_sweep.md 2> $O/r2_s9_tp.err; tail -16 $O/r2_time_profile_N_sweep.md
