#!/bin/bash
# round 2, GPU session 14: 2-D kernel -- N = 5 at 8 warps x 2 CTAs (default) vs 4 x 4, barriers between the quadrature's stages; ncu of the default
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 -x --timeout 300 > $O/r2_s14_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s14_pytest.log
tail -3 $O/r2_s14_pytest.log
MFS_B200_LIB=$PWD/ab/libmfs_inner.so timeout 600 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu -x --timeout 300 > $O/r2_s14_pytest_inner.log 2>&1
echo "inner pytest exit $?"; tail -1 $O/r2_s14_pytest_inner.log
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50" "3 18944 50" "6 4736 20" "7 2368 20"; do
  echo -n "[default] " >> $O/r2_s14_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s14_nd_ab.log 2>&1
  for v in inner w4mb4; do
    echo -n "[$v] " >> $O/r2_s14_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s14_nd_ab.log 2>&1
  done
done
done
cat $O/r2_s14_nd_ab.log
timeout 600 python tools/nd_profile_case.py 5 4736 20 > $O/r2_s14_nd_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N5_v9 -f python tools/nd_profile_case.py 5 4736 20 > $O/r2_s14_ncu.log 2>&1
cat $O/r2_s14_nd_case.log
