#!/usr/bin/env bash
# Round 2, call 5 (1 GPU): packed-arithmetic warp kernels + device solution cost: GPU suite, bench sweep, ncu.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call5
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 6 | cut -c1-700 | tee -a $O.txt; }
run timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e"
run timeout 200 $B
run env PYDCOP_B200_F2V=pipe timeout 200 $B
run env PYDCOP_B200_V2F=pipe timeout 200 $B
for cfg in "3 2 2 3" "3 2 3 2" "2 3 2 3" "2 4 2 2" "3 3 2 2" "2 3 3 2"; do
  set -- $cfg
  run env PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 timeout 200 $B
done
for cfg in "3 4 3 4" "2 6 2 7" "2 4 2 5" "3 3 3 3"; do
  set -- $cfg
  run env PYDCOP_B200_SERIAL=1 PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 timeout 200 $B
done
timeout 300 env PYDCOP_B200_SERIAL=1 PYDCOP_B200_F2VW_CPS=4 PYDCOP_B200_V2FW_CPS=4 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 12 --csv \
    --log-file ${O}_launches_serial_max.csv python bench.py --steps 5 --warmup 3 --profile > ${O}_ncu1b.log 2>&1
timeout 400 env PYDCOP_B200_V2FW_CPS=4 ncu --set full --clock-control none --import-source on -k regex:k_v2f_warp -s 2 -c 1 \
    -o ${O}_v2f_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu2.log 2>&1
timeout 400 env PYDCOP_B200_F2VW_CPS=4 ncu --set full --clock-control none --import-source on -k regex:k_f2v_warp -s 2 -c 1 \
    -o ${O}_f2v_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu3.log 2>&1
echo "== done" | tee -a $O.txt
