mkdir -p gpurun_out
TAL_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 32 --warmup 3 --no-subrecords > gpurun_out/r02_breakdown_g8.json 2> gpurun_out/r02_breakdown_g8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_breakdown_g8.json").read().strip().splitlines()[-1])
    print(8, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_us"))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_breakdown_g8.err").read()[-1500:])
PY
TAL_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 32 --warmup 3 --no-subrecords > gpurun_out/r02_breakdown_g2.json 2> gpurun_out/r02_breakdown_g2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_breakdown_g2.json").read().strip().splitlines()[-1])
    print(2, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_us"))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_breakdown_g2.err").read()[-1500:])
PY
TAL_BREAKDOWN=1 timeout 400 python bench.py --steps 32 --warmup 3 --no-subrecords --no-both --no-cpu > gpurun_out/r02_breakdown_g1.json 2> gpurun_out/r02_breakdown_g1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_breakdown_g1.json").read().strip().splitlines()[-1])
    print(1, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_us"))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_breakdown_g1.err").read()[-1500:])
PY
