#!/bin/bash
# round 2, GPU session 42: fused update/prediction for raw Benes (-DMFS_FUSE_RAW_BENES, 6 % fewer FP64 instructions) re-measured at 6 CTAs per SM
set -u
O=gpurun_out
MFS_B200_LIB=$PWD/ab/libmfs_fuse.so timeout 900 python -m pytest tests/test_gpu_filter1d.py -q -m gpu --maxfail=5 --timeout 300 > $O/r2_s42_pytest_fuse.log 2>&1; echo "fused pytest exit $?"; tail -2 $O/r2_s42_pytest_fuse.log
for rep in 1 2; do
timeout 600 python tools/ab_cases.py shipped --quick >> $O/r2_s42_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_fuse.so timeout 600 python tools/ab_cases.py fused --quick >> $O/r2_s42_ab.log 2>&1
done
timeout 300 python tools/occupancy_probe.py "shipped" 6 7 8 >> $O/r2_s42_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_fuse.so timeout 300 python tools/occupancy_probe.py "fused" 6 7 8 >> $O/r2_s42_ab.log 2>&1
cat $O/r2_s42_ab.log
