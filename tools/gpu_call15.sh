mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sharded.py tests/test_gpu_guards.py tests/test_gpu_replay.py -q -m gpu > gpurun_out/r2_t15.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2_t15.log | head -20
for OV in 1 0; do
TAL_OVERLAP=$OV timeout 600 python bench.py --steps 64 --warmup 3 --no-both --no-cpu --no-subrecords > gpurun_out/r2_bench15_ov$OV.json 2> gpurun_out/r2_bench15_ov$OV.err; echo "bench ov$OV rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench15_ov$OV.json").read().strip().splitlines()[-1])
    print("overlap $OV", {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"])
except Exception as e: print("bench parse", e); print(open("gpurun_out/r2_bench15_ov$OV.err").read()[-800:])
PY
done
