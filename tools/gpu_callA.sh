# round 2, final 1-GPU validation: full GPU suite, default bench line, C4 (VTL-PM fused / unfused), profile refresh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_pytest_gpuA.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Timeout" gpurun_out/r2_pytest_gpuA.log | head -30
for FV in 1 0; do
TAL_FUSED_VTL=$FV timeout 600 python bench.py --workload c4 --rules VTL-PM --steps 8 --warmup 2 > gpurun_out/r2_benchA_c4_fv$FV.json 2> gpurun_out/r2_benchA_c4_fv$FV.err; echo "c4 fv$FV rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_benchA_c4_fv$FV.json").read().strip().splitlines()[-1])
    r=d["rules"]["VTL-PM"]; print("c4 VTL-PM fused=$FV", r["value"], r["ms_per_step"], r["roofline"]["ms_per_launch"], r["roofline"]["frac"])
except Exception as e: print("parse", e); print(open("gpurun_out/r2_benchA_c4_fv$FV.err").read()[-1200:])
PY
done
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_benchA.json 2> gpurun_out/r2_benchA.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_benchA.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_benchA.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    c4=d.get("c4",{}); print("c4", json.dumps({k:(v if k!="rules" else {r:{kk:vv for kk,vv in x.items() if kk in ("value","ms_per_step","gpu_launches")} for r,x in v.items()}) for k,v in c4.items() if k!="config"})[:800])
    r=d.get("value_recompute",{}); print("recompute", r.get("value"), r.get("ms_per_step"))
    for k in d:
        if k.startswith("value_") and isinstance(d[k],dict) and "value" in d[k]: print(k, round(d[k]["value"],1), "e2e", d[k].get("e2e"))
except Exception as e: print("bench parse", e)
PY
timeout 300 python tools/small_configs.py 2>&1 | tail -3
bash tools/profile_r2.sh 2>&1 | tail -8
