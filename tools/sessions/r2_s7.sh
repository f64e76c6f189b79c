#!/bin/bash
# round 2, GPU session 7: shortened QL dependency chain (8-deep default, 7-deep variant, round-1 11-deep) -- parity + same-box A/B
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=8 > $O/r2_s7_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s7_pytest.log
tail -6 $O/r2_s7_pytest.log
for rep in 1 2; do
MFS_B200_LIB=$PWD/ab/libmfs_long.so timeout 600 python tools/ab_cases.py chain11 --quick >> $O/r2_s7_ab.log 2>&1
timeout 600 python tools/ab_cases.py chain8 --quick >> $O/r2_s7_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_chain7.so timeout 600 python tools/ab_cases.py chain7 --quick >> $O/r2_s7_ab.log 2>&1
done
timeout 600 python tools/ab_cases.py chain8 >> $O/r2_s7_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_chain7.so timeout 600 python tools/ab_cases.py chain7 >> $O/r2_s7_ab.log 2>&1
cat $O/r2_s7_ab.log
timeout 600 python tools/exactness_report.py > $O/r2_exactness_report.md 2> $O/r2_s7_exact.err; tail -16 $O/r2_exactness_report.md
MFS_B200_LIB=$PWD/ab/libmfs_chain7.so timeout 600 python tools/exactness_report.py > $O/r2_exactness_report_chain7.md 2>> $O/r2_s7_exact.err; tail -16 $O/r2_exactness_report_chain7.md
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s7_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v10 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s7_ncu1.log 2>&1
tail -2 $O/r2_s7_profile_case.log
for a in "5 4736 20" "5 18944 50" "4 18944 50" "7 1184 10"; do timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s7_nd.log 2>&1; done; cat $O/r2_s7_nd.log
