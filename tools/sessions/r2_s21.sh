#!/bin/bash
# round 2, GPU session 21: ncu captures of the kernels that changed last (Normal-family value kernel and gradient kernel with the step barrier, 2-D N = 7), launch list of the bench
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/grad_probe.py 7 37888 > $O/r2_s21_grad_case.log 2>&1 || { tail -5 $O/r2_s21_grad_case.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N7_normal_v12 -f python tools/grad_probe.py 7 37888 > $O/r2_s21_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_grad_kernel -c 1 -o $O/r2_filter1d_grad_N7_v12 -f python tools/grad_probe.py 7 37888 > $O/r2_s21_ncu2.log 2>&1
timeout 600 python tools/nd_profile_case.py 7 1184 10 > $O/r2_s21_nd7_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_nd_kernel -c 1 -o $O/r2_filter_nd_N7_v9 -f python tools/nd_profile_case.py 7 1184 10 > $O/r2_s21_ncu3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_s21_ncu_bench.log 2>&1
ls -la $O/*.ncu-rep | tail -4
