for ST in lower full; do
  TAL_EXCHANGE=peer timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py $ST incremental 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
done
