#!/usr/bin/env bash
# Round 2, call 11 (1 GPU): runtime-dimension factor kernel (padded rows fixed), DSA active-row kernel A/B,
# random-gather ceiling, mixed-shape side workload.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call11
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 6 | cut -c1-3000 | tee -a $O.txt; }
run timeout 900 python -m pytest tests/test_gpu_tiled_rt.py tests/test_gpu_parity.py tests/test_gpu_adsa.py tests/test_gpu_dsa_cached.py tests/test_gpu_fast.py tests/test_gpu_zz_sharded_dsa.py tests/test_gpu_fullsize.py tests/test_gpu_solve.py tests/test_gpu_dropin.py -q -p no:cacheprovider
run timeout 300 python bench.py --workload c4 --steps 100 --warmup 5
run env PYDCOP_B200_DSA_CACHE=0 timeout 300 python bench.py --workload c4 --steps 100 --warmup 5
run timeout 300 python bench.py --workload c4 --steps 20 --warmup 3
run timeout 300 python bench.py --workload mixed --steps 50 --warmup 5
run env PYDCOP_B200_TILED_RT=0 timeout 300 python bench.py --workload mixed --steps 20 --warmup 5
echo "== gather peak" | tee -a $O.txt
timeout 300 python tools/gather_peak.py 2>&1 | tail -n 30 | tee -a $O.txt
echo "== ncu dsa cached" | tee -a $O.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dsa_step_cached -s 30 -c 1 -o /tmp/dsa_cached python bench.py --workload c4 --steps 30 --warmup 5 --profile > /dev/null 2>&1
python tools/ncu_summary.py /tmp/dsa_cached.ncu-rep 2>&1 | head -60 | tee -a $O.txt
echo "== ncu mixed launch list" | tee -a $O.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file /tmp/l.csv python bench.py --workload mixed --steps 5 --warmup 3 --profile > /dev/null 2>&1
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
