# round 2, GPU call 6 (1 GPU): new tests first (metric rows, dense ROI sums, host black boxes, DMMA diagonal block),
# then the full suite, the default bench and C4 VTL-PM.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metric_rows.py tests/test_gpu_lower.py tests/test_gpu_fused.py tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/r2_t6_new.log 2>&1; echo "new tests rc=$?"
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2_t6_new.log | head -40
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu6.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu6.log | head -40
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2_bench6.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench6.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    c4=d.get("c4",{}); print("c4", json.dumps({k:(v if k!="rules" else {r:{kk:vv for kk,vv in x.items() if kk in ("value","ms_per_step","gpu_launches")} for r,x in v.items()}) for k,v in c4.items() if k!="config"})[:1200])
    r=d.get("value_recompute",{}); print("recompute", r.get("value"), r.get("ms_per_step")); 
    for k in r.get("kernels",[]): print("   ", k["kernel"][:50], round(k["ms_per_launch"],3), round(k["frac"],3))
    print("at_cpu_size", d.get("value_at_cpu_sample_size"))
except Exception as e: print("bench parse", e)
PY
