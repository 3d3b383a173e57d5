# round 2, GPU call 2: fused-step + replay tests, rankb variants (rows per CTA), bench line, full GPU suite, ncu of the pass
mkdir -p gpurun_out
for f in tests/test_gpu_fused.py tests/test_gpu_replay.py; do
  timeout 900 python -m pytest $f -q -m gpu > gpurun_out/r2_t_$(basename $f .py).log 2>&1; echo "$f rc=$?"
  tail -25 gpurun_out/r2_t_$(basename $f .py).log
done
rm -f gpurun_out/microbench_rankb.jsonl
MB_DTYPE=f64 timeout 600 python tools/microbench_rankb.py > gpurun_out/r2_mb_rankb.log 2>&1; echo "microbench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/microbench_rankb.jsonl"):
    d=json.loads(l); print(d["variant"], d["arith"], d["p"], round(d["ms"],3), round(d["frac"],3))
PY
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench1.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    print("c4", json.dumps(d.get("c4"))[:1500])
    for k in d:
        if k.startswith("value_"): print(k, json.dumps(d[k])[:300])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:600])
except Exception as e: print("bench parse", e)
PY
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2_pytest_gpu.log
for AR in fma two_rounding; do
  MB_CHILD=1 MB_DTYPE=f64 MB_ARITH=$AR MB_P=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rankb_kernel -s 3 -c 1 -o gpurun_out/r2_rankb16_$AR -f python tools/microbench_rankb.py > gpurun_out/r2_ncu_rankb16_$AR.log 2>&1
  ncu -i gpurun_out/r2_rankb16_$AR.ncu-rep --page raw --csv > gpurun_out/r2_rankb16_${AR}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_rankb16_$AR.ncu-rep --page details > gpurun_out/r2_rankb16_${AR}_details.txt 2>/dev/null
done
ls -la gpurun_out/r2_rankb16*
