#!/bin/bash
# round 2, 2-GPU session with the final kernels: the bench under torchrun exactly as the driver launches it (all legs)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2_s24_gpus.log
NCCL_DEBUG=VERSION timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/r2_bench_line_v5_n2.json 2> $O/r2_s24_bench.err
echo "bench exit $?"; tail -5 $O/r2_s24_bench.err; head -c 500 $O/r2_bench_line_v5_n2.json
