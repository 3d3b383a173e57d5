# round 2, GPU call 8 (1 GPU): c1 debug with the blocked-substitution inverse, the previously failing tests, full suite,
# sanitizer runs, bench.
mkdir -p gpurun_out
timeout 300 python tools/debug_c1.py 2>&1 | grep -E "step (0|5|12|24) |NaN" | head
timeout 900 python -m pytest tests/test_metric_rows.py tests/test_gpu_fused.py tests/test_gpu_lower.py -q -m gpu > gpurun_out/r2_t8_new.log 2>&1; echo "new tests rc=$?"
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2_t8_new.log | head -30
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu8.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu8.log | head -30
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench8.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    c4=d.get("c4",{}); print("c4", json.dumps({k:(v if k!="rules" else {r:{kk:vv for kk,vv in x.items() if kk in ("value","ms_per_step","gpu_launches")} for r,x in v.items()}) for k,v in c4.items() if k!="config"})[:800])
    r=d.get("value_recompute",{}); print("recompute", r.get("value"), r.get("ms_per_step")); 
    for k in r.get("kernels",[]): print("   ", k["kernel"][:50], round(k["ms_per_launch"],3), round(k["frac"],3))
    for k in d:
        if k.startswith("value_") and isinstance(d[k],dict) and "value" in d[k]: print(k, round(d[k]["value"],1), "e2e", d[k].get("e2e"))
except Exception as e: print("bench parse", e)
PY
timeout 300 python tools/small_configs.py 2>&1 | tail -6
bash tools/sanitize.sh
