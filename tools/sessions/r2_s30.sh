#!/bin/bash
# round 2, GPU session 30: segment-length sweep of the headline case; N sweep with best-of-3 timing
set -u
O=gpurun_out
timeout 600 python tools/segment_sweep.py > $O/r2_s30_segments.log 2>&1; cat $O/r2_s30_segments.log
timeout 600 python tools/time_profile.py > $O/r2_time_profile_N_sweep.md 2> $O/r2_s30_tp.err; tail -16 $O/r2_time_profile_N_sweep.md
