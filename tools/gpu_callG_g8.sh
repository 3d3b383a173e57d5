mkdir -p gpurun_out
TAL_EXCHANGE=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/mp_parity.py c5 40 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" | grep -E "mp parity|Error|Mismatch|Max " | tee gpurun_out/r02c_mp_parity_g8.log
