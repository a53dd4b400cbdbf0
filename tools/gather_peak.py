"""HBM throughput of random row gathers on this GPU (fg_selftest_gather): the DSA / MGM table access
pattern without the arithmetic — one JSON line per (row bytes, stride, rows in flight per thread)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydcop_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = torch.device("cuda", 0)
region = torch.empty(6 << 30, dtype=torch.uint8, device=dev)     # 6 GiB: the size of C4's oriented tables
region.zero_()
n_threads = 1_000_000
out = torch.empty(n_threads, dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for row_bytes, stride in ((80, 80), (80, 128), (128, 128), (64, 64), (32, 32), (256, 256)):
    for rows in (1, 2, 3, 6):
        rb = (row_bytes + 15) // 16 * 16
        ts = []
        for it in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = lib.fg_selftest_gather(C.c_void_p(region.data_ptr()), region.numel(), rb, max(stride, rb) // 16 * 16,
                                        n_threads * 6 // rows, rows, C.c_void_p(out.data_ptr()), st)
            b.record()
            assert rc == 0, rc
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        ms = sum(ts) / len(ts)
        useful = n_threads * 6 * rb
        print(json.dumps({"row_bytes": rb, "stride": max(stride, rb) // 16 * 16, "rows_in_flight_per_thread": rows,
                          "rows": n_threads * 6, "us": round(ms * 1e3, 2), "useful_GBs": round(useful / ms / 1e6, 1)}), flush=True)
