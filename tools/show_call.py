"""Print a gpurun call log compactly: bench JSON lines reduced to the numbers that matter."""
import json
import sys

for ln in open(sys.argv[1]):
    ln = ln.rstrip()
    if ln.startswith('{"metric"'):
        d = json.loads(ln)
        r = d.get("roofline", {})
        print("   ", str(d.get("config", {}).get("workload", ""))[:44], "N", d["n_gpus"], round(d["ms_per_step"] * 1e3, 2), "us frac",
              round(r.get("frac", 0), 3), "value %.3e" % d["value"], "clk", (d.get("clocks") or {}).get("sm_mhz"), "par",
              (d.get("parity") or {}).get("assignment_equals_single_gpu"))
        for k in ("breakdown", "e2e", "cpu_baseline", "cpu_baseline_reference_threadmode"):
            if d.get(k):
                print("       ", k, json.dumps(d[k])[:int(sys.argv[2]) if len(sys.argv) > 2 else 420])
        if d.get("run"):
            print("        run", json.dumps(d["run"])[:300])
    elif "Setting OMP_NUM_THREADS" in ln or ln.startswith("*****") or "NCCL version" in ln or not ln.strip():
        continue
    else:
        print(ln[:260])
