#!/bin/bash
# round 2, GPU session 22: 2-D kernel -- shared-memory layout with the weights in the Cholesky factor's region, one transpose buffer at S > 16, byte-wide gather table: 2 CTAs per SM at N = 7
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 --timeout 300 > $O/r2_s22_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $O/r2_s22_pytest.log
MFS_B200_LIB=$PWD/ab/libmfs_mb3.so timeout 600 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --timeout 300 -k "6 or N7" > $O/r2_s22_pytest_mb3.log 2>&1; echo "mb3 pytest exit $?"; tail -1 $O/r2_s22_pytest_mb3.log
for rep in 1 2; do
for a in "7 2368 20" "7 2368 20 tme" "6 4736 20" "5 18944 50" "4 18944 50" "3 18944 50"; do
  echo -n "[alias=default] " >> $O/r2_s22_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s22_nd_ab.log 2>&1
  echo -n "[round-1 layout] " >> $O/r2_s22_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_noalias.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s22_nd_ab.log 2>&1
done
echo -n "[alias, 3 CTAs at 168 regs] " >> $O/r2_s22_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_mb3.so timeout 300 python tools/nd_profile_case.py 6 4736 20 >> $O/r2_s22_nd_ab.log 2>&1
done
cat $O/r2_s22_nd_ab.log
timeout 600 python tools/nd_profile_case.py 7 2368 10 > $O/r2_s22_nd7_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N7_v10 -f python tools/nd_profile_case.py 7 2368 10 > $O/r2_s22_ncu.log 2>&1
