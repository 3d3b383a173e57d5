# round 2, final 8-GPU run: default bench line with the c4 / c5 (500 steps) sub-records, per-phase breakdown
mkdir -p gpurun_out
TAL_EXCHANGE=peer timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py c5 40 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" | grep -E "mp parity|Error|Mismatch|Max " | tee gpurun_out/r02b_mp_parity_g8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 32 --warmup 3 > gpurun_out/r02b_scale_g8.json 2> gpurun_out/r02b_scale_g8.err
echo "bench g8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02b_scale_g8.json").read().strip().splitlines()[-1])
    print(8, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), round(d["roofline"]["frac"],3), round(d["roofline"]["ms_per_launch"],3), d["config"].get("exchange"), d["clocks"], d["gpu_launches"])
    if "c4" in d:
        c=d["c4"]; print("c4", c.get("error") or {r:(x["value"], x["ms_per_step"]) for r,x in c["rules"].items()})
    if "c5" in d:
        c=d["c5"]; print("c5", c.get("error") or (c["value"], c["ms_per_step"], c["roofline"]["ms_per_launch"], c["roofline"]["frac"], [(k["kernel"][:20], k["ms_per_launch"], k["frac"]) for k in c["kernels"]]))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02b_scale_g8.err").read()[-1500:])
PY
