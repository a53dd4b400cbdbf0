#!/usr/bin/env bash
# Round 2, call 24 (1 GPU): where the MGM cycle goes — launch list and one --set full capture of the value-phase kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_mgm_ncu
: > $O.txt
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file /tmp/l.csv python bench.py --workload mgm --steps 10 --warmup 5 --profile > /dev/null 2>&1
python - <<'P' | tee -a $O.txt
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10][1:]
acc={}
for r in rows:
    k=r[4].split('(')[0].replace('void ','')[:60]
    acc.setdefault(k,[]).append(float(r[-1]))
print('== launch list, MGM on the C4 instance (steady state)')
for k,v in sorted(acc.items(), key=lambda kv:-sum(kv[1])):
    print('   %-62s n=%3d mean %.2f us' % (k, len(v), sum(v)/len(v)/1e3))
P
echo "== ncu --set full: k_mgm_gain_cached" | tee -a $O.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_mgm_gain_cached -s 20 -c 1 -o /tmp/mgm python bench.py --workload mgm --steps 20 --warmup 5 --profile > /dev/null 2>&1
python tools/ncu_summary.py /tmp/mgm.ncu-rep 2>&1 | head -30 | tee -a $O.txt
python tools/ncu_hot.py /tmp/mgm.ncu-rep 2>&1 | head -24 | tee -a $O.txt
echo "== done" | tee -a $O.txt
