#!/bin/bash
# round 2, GPU session 39: ncu of the final headline kernel (N = 8 at 6 CTAs per SM) + launch list of the bench
set -u
O=gpurun_out
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s39_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v12 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s39_ncu1.log 2>&1
tail -2 $O/r2_s39_profile_case.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_s39_ncu_bench.log 2>&1
ls -la $O/r2_filter1d_N8_v12.ncu-rep
