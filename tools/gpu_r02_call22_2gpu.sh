#!/usr/bin/env bash
# Round 2, call 22 (2 GPUs): the default multi-GPU path of the FINAL build (push kernels behind each side, chained close).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call22
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|destroy_process_group" | tail -n 6 | cut -c1-6000 | tee -a $O.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline
run timeout 300 python -m pytest tests/test_gpu_multiproc.py -q -p no:cacheprovider -k "p2p-2 or nccl-2"
echo "== done" | tee -a $O.txt
