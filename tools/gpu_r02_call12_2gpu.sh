#!/usr/bin/env bash
# Round 2, call 12 (2 GPUs): DSA active-row kernel v2 (thread-per-variable sums), peer-push tests (runs, split /
# joined, imbalanced, sharded cost), N=2 bench lines with the in-cycle device breakdown, side workloads sharded.
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_r02_call12_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call12
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 8 | cut -c1-4000 | tee -a $O.txt; }
run timeout 600 python -m pytest tests/test_gpu_dsa_cached.py tests/test_gpu_adsa.py tests/test_gpu_zz_sharded_dsa.py tests/test_gpu_fullsize.py -q -p no:cacheprovider -k "dsa or DSA or c4 or C4"
run timeout 300 python bench.py --workload c4 --steps 100 --warmup 5
run timeout 900 python -m pytest tests/test_gpu_multiproc.py -q -x -p no:cacheprovider -k "p2p"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5
run env PYDCOP_B200_PUSH_SPLIT=0 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run env PYDCOP_B200_PUSH_RUNS=0 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run timeout 400 $TR bench.py --gpus 2 --workload c3 --steps 200 --warmup 5
run timeout 600 $TR bench.py --gpus 2 --workload c4 --steps 100 --warmup 5
run timeout 400 $TR bench.py --gpus 2 --workload target --steps 50 --warmup 5
echo "== done" | tee -a $O.txt
