#!/usr/bin/env bash
# Round 2, call 7 (1 GPU): ncu of the packed warp kernels, summarised ON the box (reports stay there: 64 MiB cap).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call7
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 8 | cut -c1-700 | tee -a $O.txt; }
run timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_solve.py -q -p no:cacheprovider
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 12 --csv \
    --log-file ${O}_launches.csv python bench.py --steps 5 --warmup 3 --profile > ${O}_ncu1.log 2>&1
for k in v2f f2v; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_${k}_warp -s 2 -c 1 \
      -o /tmp/${k}_warp python bench.py --steps 3 --warmup 3 --profile > ${O}_ncu_${k}.log 2>&1
  python tools/ncu_summary.py /tmp/${k}_warp.ncu-rep > ${O}_${k}_warp.txt 2>&1
  python tools/ncu_hot.py /tmp/${k}_warp.ncu-rep 45 >> ${O}_${k}_warp.txt 2>&1
  python tools/ncu_lines.py /tmp/${k}_warp.ncu-rep 45 >> ${O}_${k}_warp.txt 2>&1
done
echo "== done" | tee -a $O.txt
