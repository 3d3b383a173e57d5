G=${1:-2}
for ST in lower full; do for CO in incremental recompute; do for EX in peer nccl; do
  TAL_EXCHANGE=$EX timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py $ST $CO 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
done; done; done
