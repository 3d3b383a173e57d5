# round 2, 2-GPU call: real multi-process check of the fused sharded step before the 8-GPU run
mkdir -p gpurun_out
LOG=gpurun_out/r02_mp_parity_g2.log; : > $LOG
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" | tail -3 | tee -a $LOG; }
echo "== 2 ranks: fused lower q=1"; TAL_EXCHANGE=peer run fused lower 1
echo "== 2 ranks: fused full q=2"; TAL_EXCHANGE=peer run fused full 2
echo "== 2 ranks: small lower recompute nccl"; TAL_EXCHANGE=nccl run small lower recompute
echo "== 2 ranks: c3 full size, 36 steps, replay oracle"; TAL_EXCHANGE=peer run c3 36
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/r02_scale_g2.json 2> gpurun_out/r02_scale_g2.err
echo "bench g2 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_g2.json").read().strip().splitlines()[-1])
    print(2, round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],3), round(d["roofline"]["ms_per_launch"],3), d["config"].get("exchange"), d["clocks"], d["gpu_launches"])
    for k in ("c4","c5"):
        if k in d: print(k, json.dumps(d[k])[:1500])
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02_scale_g2.err").read()[-2500:])
PY
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -q -m gpu 2>&1 | tail -3
