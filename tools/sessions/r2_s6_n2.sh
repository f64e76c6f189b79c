#!/bin/bash
# round 2, 2-GPU session: the bench under torchrun exactly as the driver launches it (all legs), plus the reference arm
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2_s6_gpus.log
NCCL_DEBUG=VERSION timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/r2_bench_line_n2.json 2> $O/r2_s6_bench.err
echo "bench exit $?"; tail -5 $O/r2_s6_bench.err; tail -c 6000 $O/r2_bench_line_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/r2_bench_reference_arm_n2.json 2>> $O/r2_s6_bench.err; tail -c 400 $O/r2_bench_reference_arm_n2.json
