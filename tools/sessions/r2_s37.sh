#!/bin/bash
# round 2, GPU session 37: new occupancy policy (6 CTAs up to N = 8, 5 at N = 9, 10): parity, full case list, and the step barrier flipped at N = 6..10
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_filter1d.py tests/test_gpu_bench_configs.py tests/test_gpu_quadrature.py -q -m gpu --maxfail=10 --timeout 300 > $O/r2_s37_pytest.log 2>&1
echo "pytest exit $?"; tail -2 $O/r2_s37_pytest.log
timeout 600 python tools/ab_cases.py shipped >> $O/r2_s37_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_barflip.so timeout 600 python tools/ab_cases.py barrier-flipped >> $O/r2_s37_ab.log 2>&1
timeout 300 python tools/occupancy_probe.py "shipped" 6 7 8 9 10 >> $O/r2_s37_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_barflip.so timeout 300 python tools/occupancy_probe.py "barrier on at 6-8, off at 9-10" 6 7 8 9 10 >> $O/r2_s37_ab.log 2>&1
cat $O/r2_s37_ab.log
