mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lower.py tests/test_gpu_guards.py tests/test_reference_fixtures.py tests/test_next_rows.py -q -m gpu > gpurun_out/r2_t12.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2_t12.log | head -20
timeout 600 python bench.py --workload c4 --rules VTL-PM --steps 8 --warmup 2 > gpurun_out/r2_bench12_c4.json 2> gpurun_out/r2_bench12_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench12_c4.json").read().strip().splitlines()[-1])
    r=d["rules"]["VTL-PM"]; print("c4 VTL-PM", r["value"], r["ms_per_step"], r["roofline"]["ms_per_launch"], r["roofline"]["frac"])
except Exception as e: print("parse", e); print(open("gpurun_out/r2_bench12_c4.err").read()[-1200:])
PY
TAL_BREAKDOWN=1 timeout 400 python bench.py --steps 64 --warmup 3 --no-subrecords --no-both --no-cpu > gpurun_out/r02_breakdown_g1.json 2> gpurun_out/r02_breakdown_g1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_breakdown_g1.json").read().strip().splitlines()[-1])
    print(1, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_us"))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_breakdown_g1.err").read()[-1500:])
PY
