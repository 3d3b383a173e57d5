# round 2, 8-GPU call: multi-process parity at 8 ranks (trimmed), the default bench line at 8 / 4 / 2 ranks
# (8: with the c4 and 500-step c5 sub-records), per-phase breakdown at 8.
mkdir -p gpurun_out
LOG=gpurun_out/r02_mp_parity_g8.log; : > $LOG
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" | tail -3 | tee -a $LOG; }
echo "== 8 ranks: fused lower q=1"; TAL_EXCHANGE=peer run fused lower 1
echo "== 8 ranks: fused full q=2"; TAL_EXCHANGE=peer run fused full 2
echo "== 8 ranks: c5 full size, 40 steps, replay oracle"; TAL_EXCHANGE=peer run c5 40
for G in 8 4; do
  EXTRA=""; if [ "$G" != "8" ]; then EXTRA="--no-subrecords"; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 32 --warmup 3 $EXTRA > gpurun_out/r02_scale_g$G.json 2> gpurun_out/r02_scale_g$G.err
  echo "bench g$G rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_g$G.json").read().strip().splitlines()[-1])
    print($G, round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],3), round(d["roofline"]["ms_per_launch"],3), d["config"].get("exchange"), d["clocks"], d["gpu_launches"])
    for k in ("c4","c5"):
        if k in d: print(k, json.dumps(d[k])[:1200])
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_scale_g$G.err").read()[-1500:])
PY
done
