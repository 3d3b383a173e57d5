# Round-2 profile of the default bench (c3, 1 GPU, lower-block storage, fused step, FMA arithmetic, cp.async-ring pass).
# Under gpurun, after `python bench.py` has exited 0 without ncu.  Outputs in gpurun_out/r02_*.
BENCH="python bench.py --steps 32 --warmup 3 --no-cpu --no-both --no-subrecords"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_default.csv \
    $BENCH > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
# the per-step kernels (3 fused kernels) and the deferred-update passes of the timed region
ncu --set full --clock-control none --import-source on -k regex:'rankb_|fused_' --launch-skip 30 -c 40 \
    -o gpurun_out/r02_step -f $BENCH > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_step.ncu-rep --page details > gpurun_out/r02_step_details.txt 2>/dev/null
python tools/ncu_traffic.py gpurun_out/r02_step_raw.csv gpurun_out/r02_kernel_traffic.json --workload c3 --n-gpus 1 --storage lower \
    --source "profiles/r02_step_raw.csv (ncu --set full --clock-control none, \`$BENCH\`)"
rm -f gpurun_out/r02_step.ncu-rep
