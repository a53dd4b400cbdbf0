#!/usr/bin/env bash
# Round 2, first GPU call: the WHOLE -m gpu suite without -x, then the pending A/B lines and captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 60 > gpurun_out/r02_call1_pytest.txt
tail -n 5 gpurun_out/r02_call1_pytest.txt
bash tools/gpu_validate_pending.sh > /dev/null 2>&1
tail -n 150 gpurun_out/pending_summary.txt
