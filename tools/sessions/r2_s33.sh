#!/bin/bash
# round 2, GPU session 33: step barrier re-checked at 4 CTAs per SM for N = 9..14 (none / barrier / shipped policy), Benes raw + Normal family
set -u
O=gpurun_out
timeout 600 python tools/ab_cases.py policy --sweep 2>&1 | grep -E "N=(9|1[0-4]) " >> $O/r2_s33_bar.log
MFS_B200_LIB=$PWD/ab/libmfs_bar0.so timeout 600 python tools/ab_cases.py none --sweep 2>&1 | grep -E "N=(9|1[0-4]) " >> $O/r2_s33_bar.log
MFS_B200_LIB=$PWD/ab/libmfs_bar1.so timeout 600 python tools/ab_cases.py barrier --sweep 2>&1 | grep -E "N=(9|1[0-4]) " >> $O/r2_s33_bar.log
timeout 600 python tools/ab_cases.py policy --sweep 2>&1 | grep -E "N=(9|1[0-4]) " >> $O/r2_s33_bar.log
cat $O/r2_s33_bar.log
