# usage: bash tools/mp_parity.sh G   (under gpurun --gpus G): every storage x conditioning x exchange combination of
# the kernel-by-kernel path, the fused step (q = 1, 2) over the peer exchange, and -- at 8 ranks -- BASELINE
# configs[4] at FULL SIZE against the low-rank replay oracle.  The log goes to profiles/ by hand.
G=${1:-2}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" | tail -3; }
for ST in lower full; do for CO in incremental recompute; do for EX in peer nccl; do
  TAL_EXCHANGE=$EX run small $ST $CO
done; done; done
for ST in lower full; do for Q in 1 2; do
  TAL_EXCHANGE=peer run fused $ST $Q
done; done
if [ "$G" = "8" ]; then TAL_EXCHANGE=peer run c5 40; fi
if [ "$G" = "4" ] || [ "$G" = "2" ]; then TAL_EXCHANGE=peer run c3 36; fi
