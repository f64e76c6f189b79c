#!/bin/bash
# round 2, GPU session 31: 1-D kernel N = 9..12 -- 2 / 3 (shipped) / 4 CTAs per SM (255 / 168 / 128 registers)
set -u
O=gpurun_out
for rep in 1 2; do
timeout 300 python tools/occupancy_probe.py "3 CTAs @168 = shipped" 9 10 11 12 >> $O/r2_s31_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb4.so timeout 300 python tools/occupancy_probe.py "4 CTAs @128" 9 10 11 12 >> $O/r2_s31_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb2.so timeout 300 python tools/occupancy_probe.py "2 CTAs @255" 9 10 11 12 >> $O/r2_s31_occ.log 2>&1
done
cat $O/r2_s31_occ.log
