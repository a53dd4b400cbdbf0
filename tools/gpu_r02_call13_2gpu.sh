#!/usr/bin/env bash
# Round 2, call 13 (2 GPUs): DSA active-row kernel v2 (fixed), tiled_rt v2 (lanes along the last dimension, big
# tables), fused halo (boundary rows stored by the warp kernels) vs push kernels, sharded cost with ghost values.
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_r02_call13_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call13
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 12 | cut -c1-4000 | tee -a $O.txt; }
run timeout 600 python -m pytest tests/test_gpu_dsa_cached.py tests/test_gpu_tiled_rt.py tests/test_gpu_adsa.py tests/test_gpu_zz_sharded_dsa.py -q -p no:cacheprovider
run timeout 300 python bench.py --workload c4 --steps 100 --warmup 5
run timeout 300 python bench.py --workload mixed --steps 50 --warmup 5
run timeout 900 python -m pytest tests/test_gpu_multiproc.py -q -p no:cacheprovider -k "p2p"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline
run env PYDCOP_B200_PUSH_FUSED=0 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run timeout 600 $TR bench.py --gpus 2 --workload c4 --steps 100 --warmup 5
run timeout 400 $TR bench.py --gpus 2 --workload target --steps 50 --warmup 5
echo "== done" | tee -a $O.txt
