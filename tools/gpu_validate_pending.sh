#!/usr/bin/env bash
# Everything that was written after round 1's GPU budget ran out, in one gpurun call (one GPU):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_validate_pending.sh'
# Outputs land in gpurun_out/pending_* ; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
run() { echo "== $*" | tee -a gpurun_out/pending_summary.txt; "$@" 2>&1 | tail -n 25 | tee -a gpurun_out/pending_summary.txt; }
: > gpurun_out/pending_summary.txt
# 1. parity: MGM kernels, direct solve entry, session with the vectorised tabulation, sharded DSA emulation
run timeout 300 python -m pytest tests/test_gpu_zz_mgm.py -q
run timeout 200 python -m pytest tests/test_gpu_solve.py tests/test_gpu_session.py -q
run timeout 300 python -m pytest tests/test_gpu_zz_fast_first.py -q
for u in 0 1; do run env PYDCOP_B200_FAST_FIRST=$u timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline; done
run timeout 300 python -m pytest tests/test_gpu_zz_dsa_v2.py -q
for u in 0 2 4; do run env PYDCOP_B200_DSA_V2=$u timeout 300 python bench.py --workload c4 --steps 200 --warmup 5; done
run timeout 300 python -m pytest tests/test_gpu_zz_sharded_dsa.py tests/test_gpu_zz_partition.py tests/test_gpu_sharded.py -q
# 2. first MGM number (C4 instance, 1M variables) and the DSA line for comparison
run timeout 300 python -m pytest tests/test_gpu_zz_mgm_fast.py -q
for u in 0 2 4; do run env PYDCOP_B200_MGM_FAST=$u timeout 300 python bench.py --workload mgm --steps 200 --warmup 5; done
# 3. launch list of the MGM step for profiles/
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/pending_mgm_launches.csv python bench.py --workload mgm --steps 10 --warmup 3 --profile \
    > gpurun_out/pending_mgm_ncu.log 2>&1
# 4. one full capture each of the DSA step (19 % of the roofline in round 1, never profiled) and the MGM gain kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_dsa_step -s 4 -c 1 \
    -o gpurun_out/pending_dsa_c4 python bench.py --workload c4 --steps 3 --warmup 3 --profile \
    > gpurun_out/pending_dsa_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mgm_gain -s 4 -c 1 \
    -o gpurun_out/pending_mgm_gain python bench.py --workload mgm --steps 3 --warmup 3 --profile \
    > gpurun_out/pending_mgm_gain_ncu.log 2>&1
echo "== done" | tee -a gpurun_out/pending_summary.txt
