mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py -q -m gpu > gpurun_out/r2_t_test_gpu_fused.log 2>&1; echo "fused rc=$?"
grep -E "^(FAILED|E  )|passed|failed|drift" gpurun_out/r2_t_test_gpu_fused.log | head -40
rm -f gpurun_out/microbench_rankb.jsonl
MB_VARIANTS=0,5,6,7,8 timeout 600 python tools/microbench_rankb.py > gpurun_out/r2_mb_rankb.log 2>&1; echo "microbench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/microbench_rankb.jsonl"):
    d=json.loads(l); print(d["variant"], d["dtype"][-7:], d["arith"], d["p"], round(d["ms"],3), round(d["frac"],3))
PY
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    print("c4", json.dumps(d.get("c4"))[:2500])
except Exception as e: print("bench parse", e)
PY
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_pytest_gpu.log
