#!/bin/bash
# round 2, GPU session 10: 2-D kernel -- shared deflation cascade (code size), register-QL fast path at N = 6, 7, occupancy A/B at N = 5
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 > $O/r2_s10_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s10_pytest.log
tail -6 $O/r2_s10_pytest.log
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50" "3 18944 50" "6 4736 20" "7 2368 20" "7 2368 20 tme"; do
  echo -n "[default] " >> $O/r2_s10_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s10_nd_ab.log 2>&1
  echo -n "[mb4]     " >> $O/r2_s10_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_ndmb4.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s10_nd_ab.log 2>&1
  echo -n "[fast16]  " >> $O/r2_s10_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_ndfast16.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s10_nd_ab.log 2>&1
done
done
cat $O/r2_s10_nd_ab.log
timeout 600 python tools/nd_profile_case.py 5 4736 20 > $O/r2_s10_nd_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N5_v8 -f python tools/nd_profile_case.py 5 4736 20 > $O/r2_s10_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N7_v8 -f python tools/nd_profile_case.py 7 1184 10 > $O/r2_s10_ncu7.log 2>&1
ls -la $O/*.ncu-rep | tail -3
