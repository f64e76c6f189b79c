#!/bin/bash
# round 2, GPU session 41: 2-D N = 5 at higher occupancy (7 warps x 3 CTAs @80 regs, 6 x 3 @96, 5 x 4 @96) vs 8 x 2 @128 (shipped)
set -u
O=gpurun_out
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme"; do
  echo -n "[8 x 2 @128 = shipped] " >> $O/r2_s41_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s41_nd_ab.log 2>&1
  for v in w7b3 w6b3 w5b4; do echo -n "[$v] " >> $O/r2_s41_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s41_nd_ab.log 2>&1; done
done
done
cat $O/r2_s41_nd_ab.log
