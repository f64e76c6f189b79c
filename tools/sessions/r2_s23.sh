#!/bin/bash
# round 2, GPU session 23: 2-D kernel N = 6, 7 -- one CTA of 8 warps per SM (all filters of the SM phase-aligned) vs two CTAs of 4
set -u
O=gpurun_out
mkdir -p $O
MFS_B200_LIB=$PWD/ab/libmfs_w8l.so timeout 600 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --timeout 300 -k "6 or N7" > $O/r2_s23_pytest_w8l.log 2>&1; echo "w8l pytest exit $?"; tail -1 $O/r2_s23_pytest_w8l.log
for rep in 1 2; do
for a in "7 2368 20" "7 2368 20 tme" "6 4736 20"; do
  echo -n "[4 warps x 2 CTAs = default] " >> $O/r2_s23_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s23_nd_ab.log 2>&1
  echo -n "[8 warps x 1 CTA] " >> $O/r2_s23_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_w8l.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s23_nd_ab.log 2>&1
done
done
cat $O/r2_s23_nd_ab.log
