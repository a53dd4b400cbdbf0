#!/usr/bin/env bash
# Round 2, call 23 (1 GPU): ncu evidence of the FINAL build — launch list of the bench command, one --set full capture
# of each hot kernel (summaries only; the reports stay on the box).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_final_ncu
: > $O.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_final_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
for k in k_f2v_warp k_v2f_warp; do
  echo "== ncu --set full: $k (C2, one launch, steady state)" | tee -a $O.txt
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o /tmp/$k python bench.py --steps 8 --warmup 3 --profile > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$k.ncu-rep 2>&1 | head -40 | tee -a $O.txt
  python tools/ncu_hot.py /tmp/$k.ncu-rep 2>&1 | head -24 | tee -a $O.txt
done
echo "== ncu --set full: k_dsa_step_cached (C4, cycle 35)" | tee -a $O.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_dsa_step_cached -s 30 -c 1 -o /tmp/dsa python bench.py --workload c4 --steps 30 --warmup 5 --profile > /dev/null 2>&1
python tools/ncu_summary.py /tmp/dsa.ncu-rep 2>&1 | head -40 | tee -a $O.txt
python tools/ncu_hot.py /tmp/dsa.ncu-rep 2>&1 | head -24 | tee -a $O.txt
echo "== done" | tee -a $O.txt
