#!/bin/bash
# round 2, GPU session 32: 1-D kernel N = 9..15 at 4 CTAs per SM (128 registers): parity + A/B incl. central mode and the Normal family
set -u
O=gpurun_out
MFS_B200_LIB=$PWD/ab/libmfs_mb4.so timeout 1500 python -m pytest tests/test_gpu_filter1d.py tests/test_gpu_bench_configs.py -q -m gpu --maxfail=10 --timeout 300 > $O/r2_s32_pytest_mb4.log 2>&1
echo "mb4 pytest exit $?"; tail -2 $O/r2_s32_pytest_mb4.log
for rep in 1 2; do
timeout 300 python tools/occupancy_probe.py "3 CTAs @168 = shipped" 13 14 15 >> $O/r2_s32_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb4.so timeout 300 python tools/occupancy_probe.py "4 CTAs @128" 13 14 15 >> $O/r2_s32_occ.log 2>&1
done
timeout 600 python tools/ab_cases.py shipped --sweep >> $O/r2_s32_occ.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_mb4.so timeout 600 python tools/ab_cases.py mb4 --sweep >> $O/r2_s32_occ.log 2>&1
cat $O/r2_s32_occ.log
