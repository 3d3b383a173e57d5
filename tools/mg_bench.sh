set -x
G=$1
for EX in peer nccl; do
  TAL_EXCHANGE=$EX TAL_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 30 --warmup 3 > gpurun_out/bench_g${G}_${EX}.json 2> gpurun_out/bench_g${G}_${EX}.err
  tail -c 600 gpurun_out/bench_g${G}_${EX}.err
  cat gpurun_out/bench_g${G}_${EX}.json | cut -c1-400
done
