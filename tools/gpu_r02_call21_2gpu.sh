#!/usr/bin/env bash
# Round 2, call 21 (2 GPUs): early q push on the main stream (default) vs off vs third stream, 300 timed steps each.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r02_call21_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call21
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|destroy_process_group" | tail -n 8 | cut -c1-6000 | tee -a $O.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e
run env PYDCOP_B200_PUSH_EARLY=0 timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run env PYDCOP_B200_PUSH_EARLY=2 timeout 300 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline
run timeout 600 python -m pytest tests/test_gpu_multiproc.py -q -p no:cacheprovider -k "maxsum and (p2p-2 or reinit-2 or late-q-push-2 or imbalanced-2)"
echo "== done" | tee -a $O.txt
