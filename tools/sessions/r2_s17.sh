#!/bin/bash
# round 2, GPU session 17: 2-D kernel -- re-associated rotation (7-deep chain) in the register QL vs the shipped one
set -u
O=gpurun_out
mkdir -p $O
MFS_B200_LIB=$PWD/ab/libmfs_chain7.so timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 --timeout 120 > $O/r2_s17_pytest_chain7.log 2>&1
echo "chain7 pytest exit $?"; tail -3 $O/r2_s17_pytest_chain7.log
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50" "3 18944 50" "6 4736 20" "7 2368 20"; do
  echo -n "[default] " >> $O/r2_s17_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s17_nd_ab.log 2>&1
  echo -n "[chain7] " >> $O/r2_s17_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_chain7.so timeout 120 python tools/nd_profile_case.py $a >> $O/r2_s17_nd_ab.log 2>&1
done
done
cat $O/r2_s17_nd_ab.log
