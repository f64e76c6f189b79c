#!/bin/bash
# round 2, GPU session 11: 2-D kernel -- deflation-cascade forms (inlined / guarded levels / switch) and a per-step CTA barrier, A/B
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_filter_nd.py -q -m gpu --maxfail=10 > $O/r2_s11_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s11_pytest.log
tail -4 $O/r2_s11_pytest.log
for rep in 1 2; do
for a in "5 18944 50" "5 18944 50 tme" "4 18944 50" "6 4736 20" "7 2368 20"; do
  echo -n "[default] " >> $O/r2_s11_nd_ab.log; timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s11_nd_ab.log 2>&1
  for v in ndL0 ndL2S2 ndbar ndbar2; do
    echo -n "[$v] " >> $O/r2_s11_nd_ab.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/nd_profile_case.py $a >> $O/r2_s11_nd_ab.log 2>&1
  done
done
done
cat $O/r2_s11_nd_ab.log
