# round 2, GPU call 4: full GPU suite with the cp.async-ring pass as the default, default bench line, C4 alone
# (VTL-PM + ITL-PM), round-2 profile.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2_bench3.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    print("c4", json.dumps(d.get("c4"))[:1500])
    for k in d:
        if k.startswith("value_"): print(k, json.dumps(d[k])[:400])
except Exception as e: print("bench parse", e)
PY
timeout 900 python bench.py --workload c4 --steps 8 --warmup 2 --itl-steps 2 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"
tail -c 1200 gpurun_out/r2_bench_c4.err; cut -c1-3000 gpurun_out/r2_bench_c4.json
bash tools/profile_r2.sh
