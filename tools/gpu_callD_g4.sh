mkdir -p gpurun_out
for OV in 1; do
TAL_OVERLAP=$OV TAL_EXCHANGE=peer timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$OV tools/mp_parity.py c3 48 > gpurun_out/r02_dbg_c3_g4_ov$OV.log 2>&1
echo "overlap=$OV rc=$?"
grep -v "^\*\*\*\|OMP_NUM\|^$\|^W[0-9]" gpurun_out/r02_dbg_c3_g4_ov$OV.log | grep -E "Error|error|assert|mp parity|Mismatch|mismatch|step [0-9]+|Max " | head -30
done
