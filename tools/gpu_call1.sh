# round 2, GPU call 1: new fused-step tests, rankb variants, a bench line, the full GPU suite, ncu of the pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_replay.py -x -q -m gpu > gpurun_out/r2_t_fused.log 2>&1; echo "fused tests rc=$?"
tail -15 gpurun_out/r2_t_fused.log
timeout 600 python tools/microbench_rankb.py > gpurun_out/r2_mb_rankb.log 2>&1; echo "microbench rc=$?"
tail -50 gpurun_out/r2_mb_rankb.log | cut -c1-200
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench1.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_gpu.log
