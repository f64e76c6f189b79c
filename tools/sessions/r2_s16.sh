#!/bin/bash
# round 2, GPU session 16: 2-D kernel -- CTA-batched stage D (one lane per matrix for the scalar QL, recorded sweeps applied by all warps) vs the shipped kernel
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 -x --timeout 300 > $O/r2_s16_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s16_pytest.log
tail -3 $O/r2_s16_pytest.log
MFS_B200_LIB=$PWD/ab/libmfs_batch.so timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 --timeout 120 > $O/r2_s16_pytest_batch.log 2>&1
echo "batch pytest exit $?"; tail -12 $O/r2_s16_pytest_batch.log
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50"; do
  echo -n "[default] " >> $O/r2_s16_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s16_nd_ab.log 2>&1
  echo -n "[batch] " >> $O/r2_s16_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_batch.so timeout 120 python tools/nd_profile_case.py $a >> $O/r2_s16_nd_ab.log 2>&1
done
done
cat $O/r2_s16_nd_ab.log
