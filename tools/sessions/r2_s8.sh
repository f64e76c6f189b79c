#!/bin/bash
# round 2, GPU session 8: rolled per-atom loops (code size) A/B, gradient kernel with the fast exp/log, 2-D timings, exactness report
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=8 > $O/r2_s8_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s8_pytest.log
tail -4 $O/r2_s8_pytest.log
for rep in 1 2; do
timeout 600 python tools/ab_cases.py unrolled --quick >> $O/r2_s8_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_roll.so timeout 600 python tools/ab_cases.py rolled --quick >> $O/r2_s8_ab.log 2>&1
done
MFS_B200_LIB=$PWD/ab/libmfs_roll.so timeout 600 python tools/ab_cases.py rolled >> $O/r2_s8_ab.log 2>&1
timeout 600 python tools/ab_cases.py unrolled >> $O/r2_s8_ab.log 2>&1
cat $O/r2_s8_ab.log
timeout 600 python tools/exactness_report.py > $O/r2_exactness_report.md 2> $O/r2_s8_exact.err; tail -16 $O/r2_exactness_report.md
timeout 600 python tools/grad_probe.py > $O/r2_s8_grad.log 2>&1; tail -12 $O/r2_s8_grad.log
for a in "5 4736 20" "5 18944 50" "4 18944 50" "3 18944 50" "6 4736 20" "7 1184 10"; do timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s8_nd.log 2>&1; done; cat $O/r2_s8_nd.log
MFS_B200_LIB=$PWD/ab/libmfs_roll.so timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s8_profile_case.log 2>&1 && \
MFS_B200_LIB=$PWD/ab/libmfs_roll.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v11_rolled -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s8_ncu1.log 2>&1
tail -2 $O/r2_s8_profile_case.log
timeout 600 python tools/bf_per_record_probe.py 8 10 > $O/r2_s8_bf_per_record.log 2>&1; cat $O/r2_s8_bf_per_record.log
for a in "5 4736 20" "5 18944 50" "4 18944 50"; do MFS_B200_LIB=$PWD/ab/libmfs_ndslow.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s8_nd_r1ql.log 2>&1; done; cat $O/r2_s8_nd_r1ql.log
