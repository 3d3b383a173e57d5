# round 2, GPU call 9 (1 GPU): guard-margin tests, 32-update pass (tests + microbench + bench at block 16 / 32), e2e of the
# early-fire host flow
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_guards.py tests/test_gpu_rankb.py tests/test_gpu_fused.py -q -m gpu > gpurun_out/r2_t9_new.log 2>&1; echo "new tests rc=$?"
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2_t9_new.log | head -30
rm -f gpurun_out/microbench_rankb.jsonl
MB_VARIANTS=-1,8 MB_DTYPE=f64 timeout 600 python tools/microbench_rankb.py > gpurun_out/r2_mb9.log 2>&1; echo "microbench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/microbench_rankb.jsonl"):
    d=json.loads(l); print(d["variant"], d["dtype"][-7:], d["arith"], d["p"], round(d["ms"],3), round(d["frac"],3), "per update", round(d["ms"]/d["p"],4))
PY
for B in 16 32; do
  timeout 600 python bench.py --steps 64 --warmup 3 --update-block $B --no-both --no-cpu --no-subrecords > gpurun_out/r2_bench9_b$B.json 2> gpurun_out/r2_bench9_b$B.err; echo "bench b$B rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench9_b$B.json").read().strip().splitlines()[-1])
    print("block $B", {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
except Exception as e: print("bench parse", e); print(open("gpurun_out/r2_bench9_b$B.err").read()[-800:])
PY
done
