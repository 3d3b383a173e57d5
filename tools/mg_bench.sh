# usage: bash tools/mg_bench.sh G [storage] [exchange...]
G=$1; ST=${2:-full}; shift; shift
EXS=${@:-peer}
for EX in $EXS; do
  TAL_EXCHANGE=$EX TAL_BREAKDOWN=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 30 --warmup 3 --storage $ST > gpurun_out/bench_g${G}_${EX}_${ST}.json 2> gpurun_out/bench_g${G}_${EX}_${ST}.err
  grep breakdown_us gpurun_out/bench_g${G}_${EX}_${ST}.err
  grep -v "^\*\*\*\|OMP_NUM\|breakdown_us\|^$" gpurun_out/bench_g${G}_${EX}_${ST}.err | tail -5
  cat gpurun_out/bench_g${G}_${EX}_${ST}.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['ms_per_launch'])"
done
