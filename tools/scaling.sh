# 2/4/8-GPU bench lines of the default workload (run under gpurun --gpus 8)
for G in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 30 --warmup 3 > gpurun_out/scale_g$G.json 2> gpurun_out/scale_g$G.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_g$G.json").read().strip().splitlines()[-1])
print($G, round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), round(d["roofline"]["achieved"]), round(d["roofline"]["ms_per_launch"],3), d["config"].get("exchange"), d["clocks"])
PY
done
TAL_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 30 --warmup 3 --workload c5 > gpurun_out/scale_c5_g8.json 2> gpurun_out/scale_c5_g8.err
grep breakdown gpurun_out/scale_c5_g8.err
cut -c1-250 gpurun_out/scale_c5_g8.json
