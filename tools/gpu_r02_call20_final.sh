#!/usr/bin/env bash
# Round 2, last call (1 GPU): whole GPU suite, smoke(), the default bench line and the reference arm on the final build.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_call20
: > $O.txt
run() { echo "== $*" | tee -a $O.txt; "$@" 2>&1 | tail -n 6 | cut -c1-6000 | tee -a $O.txt; }
run timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider
run timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run timeout 600 python bench.py
echo "== done" | tee -a $O.txt
