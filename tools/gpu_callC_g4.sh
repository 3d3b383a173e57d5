# round 2, 4-GPU run: multi-process parity at 4 ranks + bench line
mkdir -p gpurun_out
LOG=gpurun_out/r02_mp_parity_g4.log; : > $LOG
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" | tail -3 | tee -a $LOG; }
echo "== 4 ranks: fused lower q=1"; TAL_EXCHANGE=peer run fused lower 1
echo "== 4 ranks: fused full q=2"; TAL_EXCHANGE=peer run fused full 2
echo "== 4 ranks: c3 full size, 36 steps, replay oracle"; TAL_EXCHANGE=peer run c3 36
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 32 --warmup 3 --no-subrecords > gpurun_out/r02b_scale_g4.json 2> gpurun_out/r02b_scale_g4.err
echo "bench g4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02b_scale_g4.json").read().strip().splitlines()[-1])
    print(4, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), round(d["roofline"]["frac"],3), round(d["roofline"]["ms_per_launch"],3), d["config"].get("exchange"), d["clocks"], d["gpu_launches"])
except Exception as e:
    print("parse", e); print(open("gpurun_out/r02b_scale_g4.err").read()[-1500:])
PY
