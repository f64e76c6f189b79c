#!/bin/bash
# round 2, GPU session 35: 1-D kernel N = 3..8 -- CTAs per SM (register cap) 4 / 5 / 6 / 8 vs the shipped policy (6 up to N = 5, 5 at N = 6, 4 at N = 7, 8)
set -u
O=gpurun_out
for rep in 1 2; do
timeout 300 python tools/occupancy_probe.py "shipped" 3 4 5 6 7 8 >> $O/r2_s35_occ.log 2>&1
for mb in 4 5 6 8; do MFS_B200_LIB=$PWD/ab/libmfs_mb$mb.so timeout 300 python tools/occupancy_probe.py "$mb CTAs" 3 4 5 6 7 8 >> $O/r2_s35_occ.log 2>&1; done
done
cat $O/r2_s35_occ.log
