#!/bin/bash
# round 2, GPU session 36: 1-D kernel -- 5 / 6 CTAs per SM for N = 9..14, 7 for N = 6..9; the headline case (T = 1000, full history) at 5 / 6 / 7
set -u
O=gpurun_out
for rep in 1 2; do
timeout 300 python tools/occupancy_probe.py "shipped" 6 7 8 9 10 11 12 13 14 >> $O/r2_s36_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb5.so timeout 300 python tools/occupancy_probe.py "5 CTAs" 8 9 10 11 12 13 14 >> $O/r2_s36_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb6.so timeout 300 python tools/occupancy_probe.py "6 CTAs" 8 9 10 11 12 13 14 >> $O/r2_s36_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb7.so timeout 300 python tools/occupancy_probe.py "7 CTAs" 6 7 8 9 >> $O/r2_s36_occ.log 2>&1
done
for v in shipped mb5 mb6 mb7; do
  if [ $v = shipped ]; then timeout 600 python tools/ab_cases.py $v --quick >> $O/r2_s36_occ.log 2>&1; else MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 600 python tools/ab_cases.py $v --quick >> $O/r2_s36_occ.log 2>&1; fi
done
cat $O/r2_s36_occ.log
