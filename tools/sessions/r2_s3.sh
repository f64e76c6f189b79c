#!/bin/bash
# round 2, GPU session 3: QL with static post-chase instances -- parity, same-box A/B against the round-1 QL, ncu
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=8 > $O/r2_s3_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s3_pytest.log
tail -8 $O/r2_s3_pytest.log
MFS_B200_LIB=$PWD/ab/libmfs_qlv1.so timeout 600 python tools/ab_cases.py ql_v1 > $O/r2_s3_ab.log 2>&1
timeout 600 python tools/ab_cases.py ql_v2 >> $O/r2_s3_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_qlv1.so timeout 600 python tools/ab_cases.py ql_v1 --quick >> $O/r2_s3_ab.log 2>&1
cat $O/r2_s3_ab.log
timeout 600 python tools/exactness_report.py > $O/r2_exactness_report.md 2> $O/r2_s3_exact.err; cat $O/r2_exactness_report.md; tail -3 $O/r2_s3_exact.err
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s3_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v7 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s3_ncu1.log 2>&1
tail -2 $O/r2_s3_profile_case.log
