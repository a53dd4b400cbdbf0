#!/usr/bin/env bash
# Round 2, call 19 (4 GPUs): the N = 4 point of the weak-scaling line and the strong-scaled side workloads.
#   gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_r02_call19_4gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call19
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|destroy_process_group" | tail -n 6 | cut -c1-6000 | tee -a $O.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
run timeout 300 $TR bench.py --gpus 4 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e
run timeout 300 $TR bench.py --gpus 4 --workload target --steps 50 --warmup 5
run timeout 300 $TR bench.py --gpus 4 --workload c3 --steps 100 --warmup 5
echo "== done" | tee -a $O.txt
