mkdir -p gpurun_out
for G in 2 1; do
  if [ "$G" = "1" ]; then
    TAL_BREAKDOWN=1 timeout 400 python bench.py --steps 32 --warmup 3 --no-subrecords --no-both --no-cpu > gpurun_out/r02_breakdown_g1.json 2> gpurun_out/r02_breakdown_g1.err
  else
    TAL_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2952$G bench.py --gpus $G --steps 32 --warmup 3 --no-subrecords > gpurun_out/r02_breakdown_g$G.json 2> gpurun_out/r02_breakdown_g$G.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_breakdown_g$G.json").read().strip().splitlines()[-1])
    print($G, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("breakdown_us"))
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_breakdown_g$G.err").read()[-1500:])
PY
done
