mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sharded.py -q -m gpu --timeout 200 > gpurun_out/r2_tF.log 2>&1; echo "tests rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Timeout" gpurun_out/r2_tF.log | head -20
