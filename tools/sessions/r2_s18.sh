#!/bin/bash
# round 2, GPU session 18: 1-D kernel -- one CTA barrier per time step (instruction-fetch alignment, the 2-D kernel's lever) A/B
set -u
O=gpurun_out
mkdir -p $O
MFS_B200_LIB=$PWD/ab/libmfs_bar1d.so timeout 1500 python -m pytest tests/test_gpu_filter1d.py tests/test_gpu_bench_configs.py -q -m gpu --maxfail=10 --timeout 300 > $O/r2_s18_pytest_bar.log 2>&1
echo "barrier-build pytest exit $?"; tail -3 $O/r2_s18_pytest_bar.log
for rep in 1 2; do
timeout 600 python tools/ab_cases.py shipped --quick >> $O/r2_s18_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_bar1d.so timeout 600 python tools/ab_cases.py step-barrier --quick >> $O/r2_s18_ab.log 2>&1
done
MFS_B200_LIB=$PWD/ab/libmfs_bar1d.so timeout 600 python tools/ab_cases.py step-barrier >> $O/r2_s18_ab.log 2>&1
timeout 600 python tools/ab_cases.py shipped >> $O/r2_s18_ab.log 2>&1
cat $O/r2_s18_ab.log
