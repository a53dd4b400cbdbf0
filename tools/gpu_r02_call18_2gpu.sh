#!/usr/bin/env bash
# Round 2, call 18 (2 GPUs): early q push (boundary variable classes first, their rows leave on a third stream) — peer-push
# tests in every mode, N=2 bench A/B.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02_call18_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call18
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|destroy_process_group" | tail -n 10 | cut -c1-6000 | tee -a $O.txt; }
run timeout 900 python -m pytest tests/test_gpu_multiproc.py -q -p no:cacheprovider -k "p2p"
run timeout 300 python -m pytest tests/test_gpu_sharded.py -q -p no:cacheprovider
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e
run env PYDCOP_B200_PUSH_EARLY=0 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run timeout 400 $TR bench.py --gpus 2 --workload target --steps 50 --warmup 5
echo "== done" | tee -a $O.txt
