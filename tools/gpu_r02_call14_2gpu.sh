#!/usr/bin/env bash
# Round 2, call 14 (2 GPUs): N=1 regression check of the full default bench (+ reference arm), chained push (the q push
# releases the epoch and waits for the peers in its last block) vs separate release / wait kernels vs fused stores,
# push cost at 7 MB per direction (target x 2), peer-push tests, ncu of the DSA active-row kernel v2.
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_r02_call14_2gpu.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call14
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 12 | cut -c1-6000 | tee -a $O.txt; }
run timeout 600 python bench.py --steps 300 --warmup 5
run timeout 600 python bench.py --impl reference --steps 5 --warmup 3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e
run env PYDCOP_B200_PUSH_CHAIN=0 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run env PYDCOP_B200_PUSH_FUSED=1 timeout 400 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-e2e --no-cpu-baseline
run timeout 400 $TR bench.py --gpus 2 --workload target --steps 50 --warmup 5
run timeout 900 python -m pytest tests/test_gpu_multiproc.py -q -p no:cacheprovider -k "p2p and maxsum"
echo "== ncu dsa cached v2" | tee -a $O.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dsa_step_cached -s 30 -c 1 -o /tmp/dsa_cached python bench.py --workload c4 --steps 30 --warmup 5 --profile > /dev/null 2>&1
python tools/ncu_summary.py /tmp/dsa_cached.ncu-rep 2>&1 | head -40 | tee -a $O.txt
python tools/ncu_hot.py /tmp/dsa_cached.ncu-rep 2>&1 | head -40 | tee -a $O.txt
echo "== done" | tee -a $O.txt
