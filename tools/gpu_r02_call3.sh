#!/usr/bin/env bash
# Round 2, call 3 (1 GPU): the whole GPU suite with the warp-autonomous v2f kernel as default, A/B bench lines,
# launch list + full ncu capture of the new kernel, DSA fetch-granularity experiment.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call3
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 8 | cut -c1-1800 | tee -a $O.txt; }
run timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x
B="python bench.py --steps 300 --warmup 5 --no-cpu-baseline"
run timeout 200 $B
run env PYDCOP_B200_V2F=pipe timeout 200 $B
for c in 2 3 6 8; do run env PYDCOP_B200_V2FW_CPS=$c timeout 200 $B; done
run env PYDCOP_B200_SERIAL=1 timeout 200 $B
run env PYDCOP_B200_F2V_CPS=3 timeout 200 $B
run env PYDCOP_B200_F2V_CPS=4 PYDCOP_B200_V2FW_CPS=3 timeout 200 $B
for g in 32 128; do run env PYDCOP_B200_L2_FETCH=$g timeout 300 python bench.py --workload c4 --steps 100 --warmup 5; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv \
    --log-file ${O}_launches.csv python bench.py --steps 10 --warmup 3 --profile > ${O}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_v2f_warp -s 6 -c 1 \
    -o ${O}_v2f_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu2.log 2>&1
echo "== done" | tee -a $O.txt
