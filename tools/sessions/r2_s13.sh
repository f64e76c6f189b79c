#!/bin/bash
# round 2, GPU session 13: 2-D kernel -- warps (filters) per CTA with the phase barriers: 4 (default) / 6 / 8 / 12 / 16
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 -x --timeout 300 > $O/r2_s13_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s13_pytest.log
tail -3 $O/r2_s13_pytest.log
for v in w6 w12 w16; do
  MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 600 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu -x --timeout 300 -k "failed_and_missing or N5" > $O/r2_s13_pytest_$v.log 2>&1
  echo "$v pytest exit $?"; tail -1 $O/r2_s13_pytest_$v.log
done
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50" "3 18944 50"; do
  echo -n "[w4=default] " >> $O/r2_s13_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s13_nd_ab.log 2>&1
  for v in w6 w8 w12 w16; do
    echo -n "[$v] " >> $O/r2_s13_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s13_nd_ab.log 2>&1
  done
done
done
cat $O/r2_s13_nd_ab.log
