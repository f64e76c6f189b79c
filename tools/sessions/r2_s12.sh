#!/bin/bash
# round 2, GPU session 12: 2-D kernel -- counted phase barriers (0 / 2 / 3 / 4 per step), inlined cascade at every N; parity incl. failing / missing warps
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 -x --timeout 300 > $O/r2_s12_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s12_pytest.log
tail -15 $O/r2_s12_pytest.log
grep -q "pytest exit 0" $O/r2_s12_pytest.log || exit 1
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50" "3 18944 50" "2 18944 50" "6 4736 20" "7 2368 20" "7 2368 20 tme"; do
  echo -n "[bar2=default] " >> $O/r2_s12_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s12_nd_ab.log 2>&1
  for v in ndbar0 ndbar3 ndbar4; do
    echo -n "[$v] " >> $O/r2_s12_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s12_nd_ab.log 2>&1
  done
done
done
cat $O/r2_s12_nd_ab.log
