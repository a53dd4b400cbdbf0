#!/usr/bin/env bash
# Round 2, call 8 (1 GPU): v2f contiguous tiles, f2v lane remap + output staging: suite, sweep, ncu (summarised on the box).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call8
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 8 | cut -c1-600 | tee -a $O.txt; }
run timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e"
run timeout 200 $B
for cfg in "2 3 2 3" "2 3 2 4" "2 4 2 3" "2 2 2 4" "3 2 2 3" "2 3 3 2"; do
  set -- $cfg
  run env PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 timeout 200 $B
done
for cfg in "2 5 2 6" "3 3 3 4"; do
  set -- $cfg
  run env PYDCOP_B200_SERIAL=1 PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 timeout 200 $B
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 12 --csv \
    --log-file ${O}_launches.csv python bench.py --steps 5 --warmup 3 --profile > ${O}_ncu1.log 2>&1
for k in v2f f2v; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_${k}_warp -s 2 -c 1 \
      -o /tmp/${k}_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu_${k}.log 2>&1
  python tools/ncu_summary.py /tmp/${k}_warp.ncu-rep > ${O}_${k}_warp.txt 2>&1
  python tools/ncu_hot.py /tmp/${k}_warp.ncu-rep 40 >> ${O}_${k}_warp.txt 2>&1
  ncu -i /tmp/${k}_warp.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for i,n in enumerate(h):
    if any(t in n for t in ('lts__t_sectors','lts__throughput','l1tex__throughput','dram__throughput','lts__t_sector_hit','sm__inst_executed_pipe','smsp__inst_issued','l1tex__data_pipe','smsp__warp_issue_stalled','sm__pipe')): print(n, rows[1][i], v[i])
" >> ${O}_${k}_warp.txt 2>&1
done
echo "== done" | tee -a $O.txt
