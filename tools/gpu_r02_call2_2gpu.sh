#!/usr/bin/env bash
# Round 2, 2-GPU call: the real two-process tests (NCCL halo and peer push + device-side barrier), then the
# weak-scaling bench line at N=2 with the partition / split-push A/B, and the sharded side workloads.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02_call2_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call2
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 12 | cut -c1-2500 | tee -a $O.txt; }
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > ${O}_clocks.csv &
SMI=$!
run timeout 600 python -m pytest tests/test_gpu_multiproc.py -q -x -p no:cacheprovider
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run env PYDCOP_B200_PARTITION=auto timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5
run env PYDCOP_B200_PARTITION=auto PYDCOP_B200_PUSH_SPLIT=1 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5
run env PYDCOP_B200_PARTITION=blocks timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5
run env PYDCOP_B200_HALO=nccl timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5
run timeout 400 $TR bench.py --gpus 2 --workload c3 --steps 200 --warmup 5
run timeout 600 $TR bench.py --gpus 2 --workload c4 --steps 100 --warmup 5
kill $SMI
echo "== done" | tee -a $O.txt
