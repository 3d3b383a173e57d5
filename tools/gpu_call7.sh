mkdir -p gpurun_out
timeout 600 python tools/debug_c1.py 2>&1 | tail -70
