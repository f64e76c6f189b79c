#!/bin/bash
# round 2, GPU session 29: 2-D N = 5 at 7 warps x 2 CTAs (144 registers) vs 8 x 2 (128 registers, shipped)
set -u
O=gpurun_out
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme"; do
  echo -n "[8 x 2 @128 = default] " >> $O/r2_s29_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s29_nd_ab.log 2>&1
  echo -n "[7 x 2 @144] " >> $O/r2_s29_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_w7.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s29_nd_ab.log 2>&1
done
done
cat $O/r2_s29_nd_ab.log
