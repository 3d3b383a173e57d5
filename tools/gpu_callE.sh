# round 2, last 1-GPU validation of the committed state: full GPU suite + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_pytest_gpuE.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Timeout" gpurun_out/r2_pytest_gpuE.log | head -30
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_benchE.json 2> gpurun_out/r2_benchE.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_benchE.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_benchE.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    c4=d.get("c4",{}); print("c4", json.dumps({k:(v if k!="rules" else {r:{kk:vv for kk,vv in x.items() if kk in ("value","ms_per_step","gpu_launches")} for r,x in v.items()}) for k,v in c4.items() if k!="config"})[:600])
    for k in d:
        if k.startswith("value_") and isinstance(d[k],dict) and "value" in d[k]: print(k, round(d[k]["value"],1), "e2e", d[k].get("e2e"))
except Exception as e: print("bench parse", e)
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
