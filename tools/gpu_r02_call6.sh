#!/usr/bin/env bash
# Round 2, call 6 (1 GPU): GPU suite after the FFMA2 fix, ncu of the packed warp kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call6
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 8 | cut -c1-700 | tee -a $O.txt; }
run timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e"
run timeout 200 $B
run env PYDCOP_B200_SERIAL=1 timeout 200 $B
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 12 --csv \
    --log-file ${O}_launches.csv python bench.py --steps 5 --warmup 3 --profile > ${O}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_v2f_warp -s 2 -c 1 \
    -o ${O}_v2f_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_f2v_warp -s 2 -c 1 \
    -o ${O}_f2v_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu3.log 2>&1
ls -la gpurun_out | tee -a $O.txt
echo "== done" | tee -a $O.txt
