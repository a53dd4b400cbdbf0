#!/usr/bin/env bash
# Round 2, call 17 (1 GPU): MGM active-row kernel, tiled_rt hybrid mapping (small tables: one lane per output), whole suite.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call17
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 8 | cut -c1-6000 | tee -a $O.txt; }
run timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider
run timeout 300 python bench.py --workload mgm --steps 100 --warmup 5
run env PYDCOP_B200_MGM_CACHE=0 timeout 300 python bench.py --workload mgm --steps 30 --warmup 5
run timeout 300 python bench.py --workload mixed --steps 50 --warmup 5
run timeout 300 python bench.py --workload c5 --steps 200 --warmup 5
run timeout 300 python bench.py --workload c3 --steps 100 --warmup 5
run timeout 300 python bench.py --workload target --steps 50 --warmup 5
echo "== ncu mixed launch list" | tee -a $O.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file /tmp/l.csv python bench.py --workload mixed --steps 5 --warmup 3 --profile > /dev/null 2>&1
python - <<'P' | tee -a $O.txt
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10][1:]
acc={}
for r in rows:
    k=r[4].split('(')[0].replace('void ','')[:60]
    acc.setdefault(k,[]).append(float(r[-1]))
for k,v in sorted(acc.items(), key=lambda kv:-sum(kv[1])):
    print('   %-62s n=%3d mean %.2f us  sum %.1f us' % (k, len(v), sum(v)/len(v)/1e3, sum(v)/1e3))
P
echo "== done" | tee -a $O.txt
