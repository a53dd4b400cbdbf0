#!/usr/bin/env bash
# Round 2, call 9 (1 GPU): v2f with arithmetic tile descriptors; per-kernel times at several occupancies (serial launch
# lists), concurrent sweep, DSA / MGM with padded rows.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call9
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 6 | cut -c1-500 | tee -a $O.txt; }
run timeout 900 python -m pytest tests/test_gpu_tiled_rt.py tests/test_gpu_warp_kernels.py tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_zz_mgm.py tests/test_gpu_zz_mgm_fast.py tests/test_gpu_zz_sharded_dsa.py tests/test_gpu_fullsize.py tests/test_gpu_adsa.py tests/test_gpu_solve.py -q -p no:cacheprovider
B="python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e"
run timeout 200 $B
for cfg in "2 3 2 4" "2 2 2 4" "3 2 2 3" "2 3 2 3" "2 2 2 5"; do
  set -- $cfg
  run env PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 timeout 200 $B
done
for cfg in "2 5 2 6" "2 4 2 5" "3 3 2 6" "2 5 3 4"; do
  set -- $cfg
  run env PYDCOP_B200_SERIAL=1 PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 timeout 200 $B
  timeout 200 env PYDCOP_B200_SERIAL=1 PYDCOP_B200_F2VW_NS=$1 PYDCOP_B200_F2VW_CPS=$2 PYDCOP_B200_V2FW_NS=$3 PYDCOP_B200_V2FW_CPS=$4 \
    ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 8 --csv --log-file /tmp/l.csv python bench.py --steps 5 --warmup 3 --profile > /dev/null 2>&1
  python - <<'P' | tee -a $O.txt
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10][1:]
acc={}
for r in rows:
    k=r[4].split('<')[0].replace('void ','')
    acc.setdefault(k,[]).append(float(r[-1]))
print('   per-kernel us:', {k: round(sum(v)/len(v)/1e3,2) for k,v in acc.items()})
P
done
run timeout 300 python bench.py --workload mixed --steps 50 --warmup 5
run env PYDCOP_B200_TILED_RT=0 timeout 300 python bench.py --workload mixed --steps 20 --warmup 5
run timeout 300 python bench.py --workload c4 --steps 100 --warmup 5
run timeout 300 python bench.py --workload mgm --steps 100 --warmup 5
run timeout 300 python bench.py --workload c3 --steps 100 --warmup 5
run timeout 300 python bench.py --workload c5 --steps 200 --warmup 5
run timeout 300 python bench.py --workload target --steps 50 --warmup 5
echo "== done" | tee -a $O.txt
