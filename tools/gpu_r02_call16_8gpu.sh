#!/usr/bin/env bash
# Round 2, call 16 (8 GPUs): the weak-scaling line of the driver (C2 x 8) in the default mode and with the push
# kernels, the north-star instance, the grid (configs[2]) and DSA C4 (configs[3]) strong-scaled over 8 GPUs.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r02_call16_8gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call16
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|destroy_process_group" | tail -n 6 | cut -c1-6000 | tee -a $O.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
run timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e
run env PYDCOP_B200_PUSH_FUSED=0 timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e
run timeout 300 $TR bench.py --gpus 8 --workload target --steps 50 --warmup 5
run timeout 300 $TR bench.py --gpus 8 --workload c3 --steps 100 --warmup 5
run timeout 400 $TR bench.py --gpus 8 --workload c4 --steps 50 --warmup 5
echo "== done" | tee -a $O.txt
